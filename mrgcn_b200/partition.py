"""1-D node partitioning of the R-GCN stack across the GPUs of one box (new functionality —
the reference is single-process).

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  Node ranges are contiguous and
balanced by stored entries (in-edges + out-edges per node).  Rank p owns rows [lo_p, hi_p) of every
activation, the slice weight_I[:, lo_p:hi_p, :] of the identity table and the rows X[lo_p:hi_p] of the
feature matrix:

  layer 0       : source-partitioned.  Both of its terms are indexed by the SOURCE node - weight_I directly, the
                  feature term through the per-basis projection P[j] = X[j] . V (csrc/feat_proj.cu) - so rank p
                  computes the partial sums of its sources for ALL destinations and a reduce-scatter hands every
                  rank the rows it owns.  weight_I (AM: 2.67 GB) and X (1.0 GB) are never gathered or replicated;
                  backward all-gathers the pre-activation gradient (N x out), weight_I.grad stays on its shard.
                  (Without basis decomposition or with shapes the projection does not cover, the feature term of
                  layer 0 falls back to the destination-partitioned form below and needs X replicated.)
  deeper layers : destination-partitioned.  Rank p keeps the edges whose destination it owns and all-gathers H
                  (N x d floats) before the layer; backward reduce-scatters dH.
  small weights : comp / weight_F / b / relations are replicated: broadcast from rank 0 when the graph is set,
                  their gradients summed with one all-reduce (`sync_grads`).

Node numbering: every rank's activations live in a PADDED global layout of P * maxrows rows (maxrows = largest
range, rounded up to 4): node j of rank p sits at row p * maxrows + (j - lo_p).  The graphs of a rank are built
in that numbering, so that the four collectives of a step are the native equal-size, in-place
all_gather_into_tensor / reduce_scatter_tensor over (P * maxrows, d) buffers: no zero-fill, no concatenation.
The reduce-scatter of dH runs on a side stream under the weight-gradient kernels of the same layer.
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


def equal_bounds(num_nodes, parts, align=4):
    """Contiguous ranges of (nearly) equal node count - what a model built before it has seen the graph can use
    (MRGCN under torchrun, where the optimizer is created from the parameters before the first forward)."""
    return balanced_bounds(np.ones(num_nodes), parts, align)


def node_weights(row, col, num_nodes):
    """Per-node cost = entries in the node's row (feature term work) + entries whose source it is (identity
    term work).  row, col: COO index tensors of the stacked adjacency (col = rel*N + src)."""
    src = col % num_nodes
    return (torch.bincount(row, minlength=num_nodes) + torch.bincount(src, minlength=num_nodes)).cpu().numpy()


class _Layout:
    """Ranges of a partition and the padded global numbering used by the collectives."""

    def __init__(self, bounds, rank, group=None):
        self.bounds = [int(b) for b in bounds]
        self.P = len(self.bounds) - 1
        self.rank = rank
        self.group = group
        self.sizes = [self.bounds[p + 1] - self.bounds[p] for p in range(self.P)]
        self.maxrows = (max(self.sizes + [1]) + 3) // 4 * 4
        self.N = self.bounds[-1]
        self.NP = self.P * self.maxrows                      # rows of the padded global layout
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_own = self.hi - self.lo
        self.own = slice(rank * self.maxrows, rank * self.maxrows + self.n_own)
        self.side = None                                     # side stream for overlapped collectives (CUDA only)

    def pad_ids(self, ids):
        """Padded-layout row of every node id in `ids` (int64 tensor)."""
        b = torch.as_tensor(self.bounds, device=ids.device)
        p = torch.searchsorted(b, ids, right=True) - 1
        return p * self.maxrows + (ids - b[p])

    def to_padded(self, x):
        """All N rows in true node order -> the padded layout (P*maxrows rows, zeros in the padding)."""
        out = torch.zeros((self.NP,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        for p in range(self.P):
            out[p * self.maxrows:p * self.maxrows + self.sizes[p]] = x[self.bounds[p]:self.bounds[p + 1]]
        return out

    def side_stream(self, device):
        if self.side is None and device.type == "cuda":
            self.side = torch.cuda.Stream(device=device)
        return self.side


# ---- collectives over the padded layout, as autograd functions -------------------------------------------
# Communication log for bench.py's `collectives` table: (name, bytes) per call when enabled.
COMM_LOG = None


def _log(name, t):
    if COMM_LOG is not None:
        COMM_LOG.append((name, t.numel() * t.element_size()))


def _all_gather_padded(x, lay):
    """own rows (n_own, d) -> padded layout (P*maxrows, d), in place: the input of the collective is the rank's own slice of
    the output buffer."""
    d = x.shape[1]
    buf = torch.empty((lay.NP, d), dtype=x.dtype, device=x.device)
    mine = buf[lay.rank * lay.maxrows:(lay.rank + 1) * lay.maxrows]
    mine[:lay.n_own].copy_(x)
    if lay.n_own < lay.maxrows:
        mine[lay.n_own:].zero_()
    _log("all_gather", buf)
    # NCCL gathers in place (the input is the rank's slice of the output); gloo (CPU tests) wants a separate input
    dist.all_gather_into_tensor(buf, mine if x.is_cuda else mine.clone(), group=lay.group)
    return buf


def _reduce_scatter_padded(full, lay):
    """padded layout (P*maxrows, d), per-rank partial sums -> summed own rows (n_own, d)."""
    assert full.shape[0] == lay.NP
    full = full.contiguous()
    out = torch.empty((lay.maxrows, full.shape[1]), dtype=full.dtype, device=full.device)
    _log("reduce_scatter", full)
    dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=lay.group)
    return out[:lay.n_own]


def gather_rows(x, lay):
    """All-gather row blocks (no autograd) into the padded layout."""
    return _all_gather_padded(x.contiguous(), lay)


class GatherRows(torch.autograd.Function):
    """own rows (n_own, d) -> all rows in the padded layout; backward: reduce-scatter of the gradient, issued on a side
    stream as soon as the producing kernels are done (the tensor carries their event, see _LayerFn.backward) so that it runs
    under the weight-gradient kernels that follow on the main stream."""

    @staticmethod
    def forward(ctx, x, lay):
        ctx.lay = lay
        return _all_gather_padded(x.contiguous(), lay)

    @staticmethod
    def backward(ctx, g):
        lay = ctx.lay
        side = lay.side_stream(g.device)
        ev = getattr(g, "_mrgcn_ready", None)
        if side is None or ev is None:
            return _reduce_scatter_padded(g, lay).contiguous(), None
        main = torch.cuda.current_stream(g.device)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            out = _reduce_scatter_padded(g, lay).contiguous()
            done = torch.cuda.Event()
            done.record(side)
        g.record_stream(side)
        out.record_stream(main)
        main.wait_event(done)
        return out, None


class ScatterSumRows(torch.autograd.Function):
    """per-rank partial sums over the padded layout -> summed own rows; backward: all-gather."""

    @staticmethod
    def forward(ctx, full, lay):
        ctx.lay = lay
        return _reduce_scatter_padded(full, lay).contiguous()

    @staticmethod
    def backward(ctx, g):
        return _all_gather_padded(g.contiguous(), ctx.lay), None


class GatherReplicated(torch.autograd.Function):
    """own rows -> all N rows (true node order) on every rank, for callers that want the whole output (the reference's task
    loops index model(batch) with global node ids).  Every rank then computes the SAME loss from the same logits, so the
    gradient of its own rows is simply its slice of the incoming gradient: no communication in backward."""

    @staticmethod
    def forward(ctx, x, lay):
        ctx.lay = lay
        buf = _all_gather_padded(x.contiguous(), lay)
        if all(s == lay.maxrows for s in lay.sizes):
            return buf
        return torch.cat([buf[p * lay.maxrows:p * lay.maxrows + lay.sizes[p]] for p in range(lay.P)], 0)

    @staticmethod
    def backward(ctx, g):
        lay = ctx.lay
        return g[lay.lo:lay.hi].contiguous(), None


# ---- the partitioned model --------------------------------------------------------------------------------
def split_coo(row, col, val, num_nodes, R, lay):
    """(destination-partitioned COO, source-partitioned COO) of the rank that owns nodes [lo, hi), in the padded numbering.
    dst-part: entries with lo <= row < hi, rows renumbered locally, source columns in the padded layout   -> (n_own, R*NP)
    src-part: entries with lo <= src < hi, rows in the padded layout, sources renumbered locally          -> (NP, R*n_own)"""
    lo, hi = lay.lo, lay.hi
    rel, src = torch.div(col, num_nodes, rounding_mode="floor"), col % num_nodes
    mf = (row >= lo) & (row < hi)
    feat = (row[mf] - lo, rel[mf] * lay.NP + lay.pad_ids(src[mf]), val[mf])
    mi = (src >= lo) & (src < hi)
    ident = (lay.pad_ids(row[mi]), rel[mi] * (hi - lo) + (src[mi] - lo), val[mi])
    return feat, ident


class PartitionedRGCN(nn.Module):
    """The RGCN stack (mrgcn/models/rgcn.py:11-89, full-batch path) on one rank of a node partition.

    Parameter names match RGCN (`layers.layer_k.*`, `relations`); `layers.layer_0.weight_I` holds only the
    rank's slice [:, lo:hi, :] of the reference tensor (rows b*n_p + (j - lo)).  `state_dict()` returns the
    reference layout (weight_I gathered, on every rank that calls it - all ranks must call it together) and
    `load_state_dict()` accepts it, so checkpoints are interchangeable with the unpartitioned model
    (/root/reference/mrgcn/run.py:230-236)."""

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
        n_p = self.lay.n_own
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
        # the shard of weight_I is initialised with the bound of the FULL table (xavier: fan_out = S*N, graph.py:104-112)
        l0 = self.layers["layer_0"]
        S = num_bases if num_bases > 0 else num_relations
        a = float(np.sqrt(6.0 / (modules[0][1] + S * num_nodes)))
        with torch.no_grad():
            l0.weight_I.uniform_(-a, a)
        self.gF = self.gI = None
        self._register_state_dict_hook(self._gather_state)
        self._register_load_state_dict_pre_hook(self._shard_state)

    # ---- graph ---------------------------------------------------------------------------------------------
    def set_graph(self, row, col, val):
        """Build the rank's two edge sets from the full COO (device tensors) and make the replicated parameters equal on
        every rank (broadcast from rank 0)."""
        lay = self.lay
        feat, ident = split_coo(row, col, val, self.num_nodes, self.num_relations, lay)
        self.gF = self._graph_fn(*feat, lay.n_own, self.num_relations * lay.NP, self.num_relations)
        self.gI = self._graph_fn(*ident, lay.NP, self.num_relations * lay.n_own, self.num_relations)
        self.broadcast_parameters()

    def broadcast_parameters(self):
        if self.lay.P > 1 and dist.is_initialized():
            for p in self.replicated_parameters():
                dist.broadcast(p.data, src=dist.get_global_rank(self.lay.group, 0) if self.lay.group is not None else 0,
                               group=self.lay.group)

    # ---- checkpoints in the reference layout ---------------------------------------------------------------------
    def _full_weight_I(self, shard):
        lay = self.lay
        S = shard.shape[0] // max(lay.n_own, 1) if lay.n_own else (self.num_bases if self.num_bases > 0 else self.num_relations)
        out = shard.shape[1]
        pad = torch.zeros((S, lay.maxrows, out), dtype=shard.dtype, device=shard.device)
        pad[:, :lay.n_own] = shard.view(S, lay.n_own, out)
        if lay.P == 1 or not dist.is_initialized():
            return shard
        buf = torch.empty((lay.P, S, lay.maxrows, out), dtype=shard.dtype, device=shard.device)
        dist.all_gather_into_tensor(buf.view(-1), pad.view(-1), group=lay.group)
        return torch.cat([buf[p, :, :lay.sizes[p]] for p in range(lay.P)], 1).reshape(S * lay.N, out)

    @staticmethod
    def _gather_state(module, state, prefix, local_metadata):
        key = prefix + "layers.layer_0.weight_I"
        if key in state:
            state[key] = module._full_weight_I(state[key])
        return state

    def _shard_state(self, state, prefix, *args):
        key = prefix + "layers.layer_0.weight_I"
        lay = self.lay
        own_rows = self.layers["layer_0"].weight_I.shape[0]
        if key in state and state[key].shape[0] != own_rows:
            v = state[key]
            S = v.shape[0] // self.num_nodes
            state[key] = v.view(S, self.num_nodes, -1)[:, lay.lo:lay.hi, :].reshape(S * lay.n_own, -1)

    def load_full_state(self, full_state):
        """Take this rank's share of an unpartitioned RGCN state_dict (checkpoint interchange)."""
        self.load_state_dict(dict(full_state))

    # ---- forward -------------------------------------------------------------------------------------------
    def layer0_is_source_partitioned(self):
        """True when the whole input layer (identity and feature term) is computed from the rank's own sources."""
        l0 = self.layers["layer_0"]
        if l0.featureless:
            return True
        from .layers.graph import fused_projection_pitch
        return bool(self.num_bases > 0 and fused_projection_pitch(l0.indim, self.num_bases, l0.outdim))

    def forward(self, X):
        """X: features of the rank's OWN nodes (n_own, in) when layer0_is_source_partitioned(), else of all nodes in the
        padded layout (P*maxrows, in); None when featureless.  Returns the rank's rows (n_own, out_last)."""
        lay = self.lay
        H = None
        for k, (layer, act) in enumerate(zip(self.layers.values(), self.activations.values())):
            relu = isinstance(act, nn.ReLU)
            if k == 0:
                fused = self.layer0_is_source_partitioned() and not layer.featureless
                part = self._layer_fn(X if fused else None, layer.weight_I, layer.weight_I_comp,
                                      layer.weight_F if fused else None, layer.weight_F_comp if fused else None, None, None,
                                      self.gI, self.gI if fused else None, self.num_bases, False)   # (NP, out) partial sums
                own = ScatterSumRows.apply(part, lay)                                               # (n_own, out)
                if layer.featureless or fused:
                    H = own if layer.b is None else own + layer.b
                    H = torch.relu(H) if relu else H
                else:
                    H = self._layer_fn(X, None, None, layer.weight_F, layer.weight_F_comp, layer.b, None, None, self.gF,
                                       self.num_bases, relu, own)
            else:
                Hall = GatherRows.apply(H, lay)                                                     # (NP, d)
                H = self._layer_fn(Hall, None, None, layer.weight_F, layer.weight_F_comp, layer.b, None, None, self.gF,
                                   self.num_bases, relu)
            if act is not None and not relu:
                H = act(H)
        return H

    def forward_all(self, X):
        """Logits of ALL nodes on every rank (true node order): what the reference's task loops expect from model(batch)."""
        return GatherReplicated.apply(self.forward(X), self.lay)

    def replicated_parameters(self):
        return [p for n, p in self.named_parameters() if n != "layers.layer_0.weight_I"]

    def enable_grad_hooks(self):
        """For callers that do not know about `sync_grads` (the reference's unchanged task loops): all-reduce the gradient of
        every replicated parameter as soon as autograd has accumulated it."""
        if getattr(self, "_hooked", False) or self.lay.P == 1:
            return
        self._hooked = True
        self.hooks_enabled = True      # a caller that sums the gradients itself (sync_grads: one all-reduce) switches them off

        def hook(q):
            if self.hooks_enabled:
                dist.all_reduce(q.grad, op=dist.ReduceOp.SUM, group=self.lay.group)
        for p in self.replicated_parameters():
            p.register_post_accumulate_grad_hook(hook)

    def sync_grads(self):
        """Sum the gradients of the replicated (small) parameters over ranks with ONE all-reduce."""
        ps = [p for p in self.replicated_parameters() if p.grad is not None]
        if not ps or self.lay.P == 1:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        _log("all_reduce", flat)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.lay.group)
        off = 0
        for p in ps:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
