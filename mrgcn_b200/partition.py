"""1-D node partitioning of the R-GCN stack across the GPUs of one box (new functionality —
the reference is single-process).

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  Node ranges are contiguous and
balanced by stored entries (in-edges + out-edges per node).  Rank p owns rows [lo_p, hi_p) of every
activation and the slice weight_I[:, lo_p:hi_p, :] of the identity table:

  feature term  : destination-partitioned.  Rank p keeps the edges whose destination it owns and needs the
                  features of ALL sources: layer 0 reads the (static) feature matrix, deeper layers
                  all-gather H (N x d floats) before the layer; backward reduce-scatters dH.
  identity term : source-partitioned, because weight_I is indexed by source and is far too large to gather
                  (AM: 2.67 GB vs 67 MB for an N x out activation).  Rank p computes the partial sums of its
                  sources for ALL destinations, a reduce-scatter hands every rank the rows it owns; backward
                  all-gathers the pre-activation gradient, and weight_I.grad stays local to the shard.
  small weights : comp / weight_F / b / relations are replicated; their gradients are summed with one
                  all-reduce (`sync_grads`).

Rows are padded to the largest range so that the NCCL collectives are the native equal-size
all_gather_into_tensor / reduce_scatter_tensor.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from .graph import RelGraph
from .layers.graph import GraphConvolution, _LayerFn


def balanced_bounds(weight, parts, align=4):
    """Contiguous ranges [b[p], b[p+1]) over len(weight) nodes whose weight sums are as equal as a prefix
    cut allows.  weight: 1-D array of non-negative per-node costs (stored entries touching the node).
    Interior cuts are rounded to multiples of `align` nodes so that a shard's rows of weight_I start 16-byte
    aligned (the TMA-engine kernels need that)."""
    w = np.asarray(weight, dtype=np.float64)
    n = len(w)
    c = np.concatenate([[0.0], np.cumsum(w + 1e-9)])
    targets = c[-1] * np.arange(1, parts) / parts
    cuts = np.searchsorted(c, targets, side="left")
    if align > 1 and n >= parts * align:
        cuts = (cuts + align // 2) // align * align
    b = np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(b)


def node_weights(row, col, num_nodes):
    """Per-node cost = entries in the node's row (feature term work) + entries whose source it is (identity
    term work).  row, col: COO index tensors of the stacked adjacency (col = rel*N + src)."""
    src = col % num_nodes
    return (torch.bincount(row, minlength=num_nodes) + torch.bincount(src, minlength=num_nodes)).cpu().numpy()


# ---- collectives over row blocks, as autograd functions --------------------------------------------------
class _Layout:
    def __init__(self, bounds, rank, group=None):
        self.bounds = [int(b) for b in bounds]
        self.P = len(self.bounds) - 1
        self.rank = rank
        self.group = group
        self.sizes = [self.bounds[p + 1] - self.bounds[p] for p in range(self.P)]
        self.maxrows = max(self.sizes) if self.sizes else 0
        self.N = self.bounds[-1]
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]


def _all_gather_rows(x, lay):
    d = x.shape[1]
    pad = torch.zeros((lay.maxrows, d), dtype=x.dtype, device=x.device)
    pad[:x.shape[0]] = x
    buf = torch.empty((lay.P * lay.maxrows, d), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(buf, pad, group=lay.group)
    if all(s == lay.maxrows for s in lay.sizes):
        return buf
    return torch.cat([buf[p * lay.maxrows:p * lay.maxrows + lay.sizes[p]] for p in range(lay.P)], 0)


def _reduce_scatter_rows(full, lay):
    d = full.shape[1]
    if all(s == lay.maxrows for s in lay.sizes):
        buf = full.contiguous()
    else:
        buf = torch.zeros((lay.P * lay.maxrows, d), dtype=full.dtype, device=full.device)
        for p in range(lay.P):
            buf[p * lay.maxrows:p * lay.maxrows + lay.sizes[p]] = full[lay.bounds[p]:lay.bounds[p + 1]]
    out = torch.empty((lay.maxrows, d), dtype=full.dtype, device=full.device)
    dist.reduce_scatter_tensor(out, buf, op=dist.ReduceOp.SUM, group=lay.group)
    return out[:lay.sizes[lay.rank]]


def gather_rows(x, lay):
    """All-gather row blocks (no autograd): rank p contributes rows [bounds[p], bounds[p+1])."""
    return _all_gather_rows(x.contiguous(), lay)


class GatherRows(torch.autograd.Function):
    """local rows (n_p, d) -> all rows (N, d); backward: reduce-scatter of the gradient."""

    @staticmethod
    def forward(ctx, x, lay):
        ctx.lay = lay
        return _all_gather_rows(x.contiguous(), lay)

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_rows(g.contiguous(), ctx.lay).contiguous(), None


class ScatterSumRows(torch.autograd.Function):
    """per-rank partial sums over all rows (N, d) -> summed local rows (n_p, d); backward: all-gather."""

    @staticmethod
    def forward(ctx, full, lay):
        ctx.lay = lay
        return _reduce_scatter_rows(full, lay).contiguous()

    @staticmethod
    def backward(ctx, g):
        return _all_gather_rows(g.contiguous(), ctx.lay), None


# ---- the partitioned model --------------------------------------------------------------------------------
def split_coo(row, col, val, num_nodes, R, lo, hi):
    """(feature-term COO, identity-term COO) of the rank that owns nodes [lo, hi).
    feature : entries with lo <= row < hi, rows renumbered, all N source columns     -> shape (hi-lo, R*N)
    identity: entries with lo <= src < hi, all N rows, sources renumbered             -> shape (N, R*(hi-lo))"""
    mf = (row >= lo) & (row < hi)
    feat = (row[mf] - lo, col[mf], val[mf])
    rel, src = torch.div(col, num_nodes, rounding_mode="floor"), col % num_nodes
    mi = (src >= lo) & (src < hi)
    ident = (row[mi], rel[mi] * (hi - lo) + (src[mi] - lo), val[mi])
    return feat, ident


class PartitionedRGCN(nn.Module):
    """The RGCN stack (mrgcn/models/rgcn.py:11-89, full-batch path) on one rank of a node partition.

    Parameter names match RGCN (`layers.layer_k.*`, `relations`); `layers.layer_0.weight_I` holds only the
    rank's slice [:, lo:hi, :] of the reference tensor (rows b*n_p + (j - lo))."""

    def __init__(self, modules, num_relations, num_nodes, num_bases, featureless, bias, link_prediction,
                 bounds, rank, group=None, layer_fn=None, graph_fn=None):
        """layer_fn / graph_fn: test seams (tests/test_partition_gloo.py drives the partition logic and the
        collectives on CPU with the oracle's layer arithmetic); the product default is the CUDA layer."""
        super().__init__()
        self._layer_fn = layer_fn or _LayerFn.apply
        self._graph_fn = graph_fn or RelGraph.from_coo_arrays
        self.lay = _Layout(bounds, rank, group)
        self.num_nodes, self.num_relations, self.num_bases = num_nodes, num_relations, num_bases
        self.featureless = featureless
        n_p = self.lay.hi - self.lay.lo
        self.layers = nn.ModuleDict()
        self.activations = nn.ModuleDict()
        for k, (indim, outdim, _ltype, act) in enumerate(modules):
            # the input layer is built for the local node count: its weight_I is the shard
            self.layers["layer_%d" % k] = GraphConvolution(indim, outdim, num_relations, n_p if k == 0 else num_nodes,
                                                            num_bases=num_bases, bias=bias, input_layer=(k == 0),
                                                            featureless=(featureless if k == 0 else False))
            self.activations["layer_%d" % k] = act
        self.num_layers = len(self.layers)
        if link_prediction:
            self.relations = nn.Parameter(torch.empty((num_relations, modules[-1][1])))
            nn.init.xavier_uniform_(self.relations)
        self.gF = self.gI = None

    def set_graph(self, row, col, val):
        """Build the rank's two edge sets from the full COO (device tensors)."""
        lay = self.lay
        feat, ident = split_coo(row, col, val, self.num_nodes, self.num_relations, lay.lo, lay.hi)
        n_p = lay.hi - lay.lo
        self.gF = self._graph_fn(*feat, n_p, self.num_relations * self.num_nodes, self.num_relations)
        self.gI = self._graph_fn(*ident, self.num_nodes, self.num_relations * n_p, self.num_relations)

    def load_full_state(self, full_state):
        """Take this rank's share of an unpartitioned RGCN state_dict (checkpoint interchange)."""
        lay = self.lay
        own = self.state_dict()
        for k, v in full_state.items():
            if k == "layers.layer_0.weight_I":
                S = v.shape[0] // self.num_nodes
                v = v.view(S, self.num_nodes, -1)[:, lay.lo:lay.hi, :].reshape(S * (lay.hi - lay.lo), -1)
            own[k].copy_(v)

    def forward(self, X):
        """X: (N, in) features of ALL nodes (static input, replicated) or None when featureless.
        Returns the rank's rows (n_p, out_last)."""
        lay = self.lay
        H = None
        for k, (layer, act) in enumerate(zip(self.layers.values(), self.activations.values())):
            relu = isinstance(act, nn.ReLU)
            if k == 0:
                part = self._layer_fn(None, layer.weight_I, layer.weight_I_comp, None, None, None, None, self.gI, None,
                                      self.num_bases, False)                       # (N, out) partial, my sources only
                own = ScatterSumRows.apply(part, lay)                              # (n_p, out)
                if layer.featureless:
                    H = own if layer.b is None else own + layer.b
                    H = torch.relu(H) if relu else H
                else:
                    H = self._layer_fn(X, None, None, layer.weight_F, layer.weight_F_comp, layer.b, None, None, self.gF,
                                       self.num_bases, relu, own)
            else:
                Hall = GatherRows.apply(H, lay)                                    # (N, d)
                H = self._layer_fn(Hall, None, None, layer.weight_F, layer.weight_F_comp, layer.b, None, None, self.gF,
                                   self.num_bases, relu)
            if act is not None and not relu:
                H = act(H)
        return H

    def replicated_parameters(self):
        return [p for n, p in self.named_parameters() if n != "layers.layer_0.weight_I"]

    def sync_grads(self):
        """Sum the gradients of the replicated (small) parameters over ranks with ONE all-reduce."""
        ps = [p for p in self.replicated_parameters() if p.grad is not None]
        if not ps or self.lay.P == 1:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.lay.group)
        off = 0
        for p in ps:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
