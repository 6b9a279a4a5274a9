"""R-GCN layer: same constructor, parameters, state_dict and forward signature as the reference's
`mrgcn.layers.graph.GraphConvolution` (/root/reference/mrgcn/layers/graph.py:8-116); the arithmetic
runs in hand-written sm_100a kernels behind the C ABI (include/mrgcn_b200.h).  No CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from .. import _native as nv
from ..graph import RelGraph, graph_of


def _empty(n, like_dev):
    return torch.empty(max(int(n), 1), dtype=torch.float32, device=like_dev)


def _tab_mode(B, out_dim):
    """Which table-term kernels (csrc/tab.cu) apply (bit 0 messages, 1 basis gradient, 2 comp gradient); MRGCN_TAB=<mask>
    in the environment restricts them to compare against the tile-staging kernels of round 1."""
    return int(nv.lib().mrgcn_tab_mode(B, 0, out_dim))


def padded_features(X, pitch=None):
    """Feature matrix stored with rows of ceil32(in) floats (zeros beyond `in`): the layout the tensor-map TMA loads of
    the projection kernel (csrc/feat_proj.cu) want.  Returns the (N, in) VIEW of that buffer; handing it to the layer avoids
    the per-call padding pass (the bench and MRGCN's feature upload do this)."""
    n, d = X.shape
    pitch = pitch or ((d + 31) // 32) * 32
    buf = torch.zeros((n, pitch), dtype=torch.float32, device=X.device)
    buf[:, :d].copy_(X)
    return buf[:, :d]


def _x_layout(X):
    """(data tensor handed to the kernels, row pitch in floats).  Accepts contiguous matrices and row-padded views."""
    if X.dim() == 2 and X.stride(1) == 1 and X.stride(0) >= X.shape[1] and X.storage_offset() % 4 == 0 and X.shape[0] > 1:
        return X, int(X.stride(0))
    Xc = X.contiguous()
    return Xc, int(Xc.shape[1])


def fused_projection_pitch(in_dim, B, out_dim):
    """Row pitch (floats) the projection kernel wants for X when an input layer with identity and feature term and B
    bases runs its feature term through csrc/feat_proj.cu + the table-term kernels; 0 when that path does not apply."""
    if not B or B <= 0 or os.environ.get("MRGCN_PROJ", "1") == "0" or not _tab_mode(B, out_dim):
        return 0
    pitch = int(nv.lib().mrgcn_feat_proj_supported(in_dim, B, out_dim))
    return pitch if pitch and (int(nv.lib().mrgcn_tab_mode(B, B, out_dim)) & 1) else 0


def _hub_ws(gI, gF, in_dim, out_dim, B, dev):
    """Workspace for the partial sums of hub segments (include/mrgcn_b200.h: hub_ws)."""
    segs = max([max(g.n_row_segs, g.n_col_segs) for g in (gI, gF) if g is not None] + [0])
    return _empty(segs * max(out_dim, in_dim, max(B, 1) * out_dim), dev) if segs else None


class _LayerFn(torch.autograd.Function):
    """out = act(mask * (b + A.W_I(mixed) + A.(X W_F(mixed))))  — graph.py:62-102 + rgcn.py:78-87."""

    @staticmethod
    def forward(ctx, X, weight_I, comp_I, weight_F, comp_F, bias, row_mask, gI, gF, B, relu, addend=None):
        hasI, hasF = weight_I is not None, X is not None
        ref = weight_I if hasI else weight_F
        dev = ref.device
        out_dim = ref.shape[-1]
        in_dim = weight_F.shape[1] if hasF else 0
        B = int(B) if B and B > 0 else 0
        g0 = gI if hasI else gF
        tens = dict(X=X, weight_I=weight_I, comp_I=comp_I if (hasI and B) else None,
                    weight_F=weight_F if hasF else None, comp_F=comp_F if (hasF and B) else None,
                    bias=bias, row_mask=row_mask, addend=addend)
        x_stride = 0
        for k, t in tens.items():
            if t is not None:
                nv.require_cuda(t, k)
                if t.dtype != torch.float32:
                    raise TypeError("mrgcn_b200: %s must be float32" % k)
                if k == "X":
                    tens[k], x_stride = _x_layout(t)
                else:
                    tens[k] = t.contiguous()
        if hasF and tens["X"].shape != (gF.NS, in_dim):
            raise ValueError("X has shape %s, expected (%d, %d)" % (tuple(X.shape), gF.NS, in_dim))
        if hasI and weight_I.shape[0] != (B if B else gI.R) * gI.NS:
            raise ValueError("weight_I has %d rows, expected %d" % (weight_I.shape[0], (B if B else gI.R) * gI.NS))
        out = torch.empty((g0.ND, out_dim), dtype=torch.float32, device=dev)
        wmix = _empty(gF.R * in_dim * out_dim, dev) if (hasF and B) else None
        ms = int(nv.lib().mrgcn_msg_stride(out_dim))
        # table-term kernels (csrc/tab.cu) for an input layer with basis decomposition
        plan = gI.tab_plan() if (hasI and B and _tab_mode(B, out_dim)) else None
        # ... with identity AND feature term: project the features per basis on the tensor cores (csrc/feat_proj.cu) and
        # mix both tables in one pass; no per-edge feature messages then
        proj = vt_ws = xpad_ws = None
        if plan is not None and hasF and gI is gF:
            pitch = fused_projection_pitch(in_dim, B, out_dim)
            if pitch:
                proj = _empty(gI.NS * B * out_dim, dev)
                vt_ws = _empty(2 * B * out_dim * pitch, dev)
                if x_stride != pitch:
                    xpad_ws = _empty(gI.NS * pitch, dev)
        msg_I = _empty(gI.E * ms, dev) if (hasI and B) else None
        # feature-only narrow layers run in one pass (csrc/narrow.cu): no per-edge messages either
        narrow_f = hasF and not hasI and bool(nv.lib().mrgcn_narrow_supported(gF.R, in_dim, out_dim))
        msg_F = _empty(gF.E * ms, dev) if (hasF and proj is None and not narrow_f) else None
        a = nv.LayerArgs()
        a.gI = C.pointer(gI.c) if hasI else None
        a.gF = C.pointer(gF.c) if hasF else None
        a.in_dim, a.out_dim, a.B, a.relu, a.x_stride = in_dim, out_dim, B, int(bool(relu)), x_stride
        for k, t in tens.items():
            setattr(a, k, nv.ptr(t))
        a.wmix, a.msg_I, a.msg_F, a.out = nv.ptr(wmix), nv.ptr(msg_I), nv.ptr(msg_F), nv.ptr(out)
        hub_ws = _hub_ws(gI if hasI else None, gF if hasF else None, in_dim, out_dim, B, dev)
        a.hub_ws = nv.ptr(hub_ws)
        a.plan = C.pointer(plan) if plan is not None else None
        a.proj, a.vt_ws, a.xpad_ws = nv.ptr(proj), nv.ptr(vt_ws), nv.ptr(xpad_ws)
        with torch.cuda.device(dev):
            nv.check(nv.lib().mrgcn_rgcn_layer_fwd(C.byref(a), nv.stream_ptr()), "rgcn_layer_fwd")
        ctx.gI, ctx.gF, ctx.B, ctx.relu, ctx.dims, ctx.x_stride = gI, gF, B, bool(relu), (in_dim, out_dim), x_stride
        ctx.has_addend = addend is not None
        ctx.save_for_backward(tens["X"], tens["weight_I"], tens["comp_I"], tens["weight_F"], tens["comp_F"],
                              tens["bias"], tens["row_mask"], wmix, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        X, weight_I, comp_I, weight_F, comp_F, bias, row_mask, wmix, out = ctx.saved_tensors
        gI, gF, B = ctx.gI, ctx.gF, ctx.B
        in_dim, out_dim = ctx.dims
        hasI, hasF = weight_I is not None, X is not None
        dev = out.device
        gout = gout.contiguous().float()
        need = ctx.needs_input_grad
        b = nv.LayerBwdArgs()
        f = b.f
        f.gI = C.pointer(gI.c) if hasI else None
        f.gF = C.pointer(gF.c) if hasF else None
        f.in_dim, f.out_dim, f.B, f.relu, f.x_stride = in_dim, out_dim, B, int(ctx.relu), ctx.x_stride
        f.weight_I, f.comp_I, f.X, f.weight_F, f.comp_F = (nv.ptr(weight_I), nv.ptr(comp_I), nv.ptr(X),
                                                           nv.ptr(weight_F), nv.ptr(comp_F))
        f.bias, f.row_mask, f.wmix, f.out = nv.ptr(bias), nv.ptr(row_mask), nv.ptr(wmix), nv.ptr(out)
        g0 = gI if hasI else gF
        gact = _empty(g0.ND * out_dim, dev)
        g_wI = g_cI = g_wF = g_cF = g_b = g_X = None
        cbuf = g_wmix = None
        part_elems = 1
        mode = _tab_mode(B, out_dim) if (hasI and B) else 0
        plan = gI.tab_plan() if mode else None
        f.plan = C.pointer(plan) if plan is not None else None
        if hasI and (need[1] or need[2]):
            g_wI = torch.empty_like(weight_I)
            if B:
                g_cI = torch.empty_like(comp_I)
                if mode & 4:      # records of the (tile, relation) pieces instead of an E x B scratch
                    cbuf = _empty((plan.n_pieces + plan.n_blks) * B, dev)
                else:
                    cbuf = _empty(gI.E * B, dev)
                    part_elems = max(part_elems, gI.n_chunks * B)
        if hasF and (need[3] or need[4]):
            g_wF = torch.empty_like(weight_F)
            part_elems = max(part_elems, gF.n_chunks * in_dim * out_dim)
            if B:
                g_cF = torch.empty_like(comp_F)
                g_wmix = _empty(gF.R * in_dim * out_dim, dev)
        wt_ws = msgx_ws = None
        if hasF and need[0]:
            g_X = torch.empty(X.shape, dtype=torch.float32, device=dev)
            wt_ws = _empty(gF.R * in_dim * out_dim, dev)
            if nv.lib().mrgcn_narrow_supported(gF.R, out_dim, in_dim):
                msgx_ws = _empty(4, dev)          # the input gradient runs in one pass: no per-edge messages
            else:
                msgx_ws = _empty(gF.E * int(nv.lib().mrgcn_msg_stride(in_dim)), dev)
        colsum = None
        if bias is not None and need[5]:
            g_b = torch.empty_like(bias)
            colsum = _empty(((g0.ND + 1023) // 1024) * out_dim, dev)
        part = _empty(part_elems, dev)
        b.gout = nv.ptr(gout)
        b.g_weight_I, b.g_comp_I, b.g_weight_F, b.g_comp_F = nv.ptr(g_wI), nv.ptr(g_cI), nv.ptr(g_wF), nv.ptr(g_cF)
        b.g_bias, b.g_X = nv.ptr(g_b), nv.ptr(g_X)
        b.gact, b.cbuf, b.part, b.g_wmix, b.colsum_ws = nv.ptr(gact), nv.ptr(cbuf), nv.ptr(part), nv.ptr(g_wmix), nv.ptr(colsum)
        b.wt_ws, b.msgx_ws = nv.ptr(wt_ws), nv.ptr(msgx_ws)
        hub_ws = _hub_ws(gI if hasI else None, gF if hasF else None, in_dim, out_dim, B, dev)
        f.hub_ws = nv.ptr(hub_ws)
        with torch.cuda.device(dev):
            if g_X is not None and not hasI and torch.distributed.is_available() and torch.distributed.is_initialized():
                # hidden layer of a node-partitioned run: the input gradient feeds a reduce-scatter (partition.GatherRows);
                # compute it first and mark the point where it is ready, so that the collective can run on a side stream
                # under the weight-gradient kernels launched next
                b.phases = nv.BWD_ACT | nv.BWD_GX
                nv.check(nv.lib().mrgcn_rgcn_layer_bwd(C.byref(b), nv.stream_ptr()), "rgcn_layer_bwd")
                g_X._mrgcn_ready = torch.cuda.Event()
                g_X._mrgcn_ready.record()
                b.phases = nv.BWD_IDENT | nv.BWD_FEATW
            nv.check(nv.lib().mrgcn_rgcn_layer_bwd(C.byref(b), nv.stream_ptr()), "rgcn_layer_bwd")
        g_add = gact[:g0.ND * out_dim].view(g0.ND, out_dim) if (ctx.has_addend and need[11]) else None
        return (g_X, g_wI, g_cI, g_wF, g_cF, g_b, None, None, None, None, None, g_add)


def slice_columns_device(A, A_idx, device):
    """`sliceSparseCOO` (/root/reference/mrgcn/data/batch.py:252-263) on the device: keep the entries whose
    column is in A_idx, renumber columns by position in A_idx, reset every kept value to 1.0 (float32)."""
    ind = A._indices().to(device)
    col_idx = torch.as_tensor(A_idx, dtype=torch.int64, device=device)
    order = torch.argsort(col_idx, stable=True)
    sorted_cols = col_idx[order]
    pos = torch.searchsorted(sorted_cols, ind[1]).clamp_(max=max(len(col_idx) - 1, 0))
    keep = sorted_cols[pos] == ind[1] if len(col_idx) else torch.zeros_like(ind[1], dtype=torch.bool)
    row, col = ind[0][keep], order[pos[keep]]
    return row, col, torch.ones(len(row), dtype=torch.float32, device=device)


class GraphConvolution(nn.Module):
    """Relational graph convolution layer (drop-in for mrgcn.layers.graph.GraphConvolution).

    Parameters, their shapes, registration order (weight_I_comp, weight_F_comp, weight_I, weight_F, b) and
    initialisation follow graph.py:9-60,104-116, so the same seed gives the same initial state_dict.
    """

    def __init__(self, indim, outdim, num_relations, num_nodes, num_bases=-1, bias=False, input_layer=False,
                 featureless=False, shared_bases_weights=False):
        super().__init__()
        self.indim, self.outdim = indim, outdim
        self.num_relations, self.num_nodes, self.num_bases = num_relations, num_nodes, num_bases
        self.input_layer, self.featureless, self.bias = input_layer, featureless, bias
        self.weight_I = self.weight_F = self.weight_I_comp = self.weight_F_comp = self.b = None
        S = num_relations
        if num_bases > 0:
            S = num_bases
            if input_layer:
                self.weight_I_comp = nn.Parameter(torch.empty((num_relations, num_bases)))
            if not featureless:
                if shared_bases_weights:
                    self.weight_F_comp = self.weight_I_comp
                else:
                    self.weight_F_comp = nn.Parameter(torch.empty((num_relations, num_bases)))
        if input_layer:
            self.weight_I = nn.Parameter(torch.empty((S * num_nodes, outdim)))
        if not featureless:
            self.weight_F = nn.Parameter(torch.empty((S, indim, outdim)))
        if bias:
            self.b = nn.Parameter(torch.empty(outdim))
        self.reset_parameters()

    def reset_parameters(self):
        for name, param in self.named_parameters():
            if name == "b":
                continue
            nn.init.xavier_uniform_(param)
        if self.bias:
            nn.init.zeros_(self.b)

    def _device(self):
        p = self.weight_I if self.weight_I is not None else self.weight_F
        if not p.is_cuda:
            raise RuntimeError("mrgcn_b200.GraphConvolution has no CPU path: move the module to a CUDA device "
                               "(task.gcn_gpu_acceleration = true)")
        return p.device

    def forward(self, X, A, A_idx=None, *, row_mask=None, relu=False, addend=None):
        """Same contract as graph.py:62: X None (featureless input layer) or (n, indim) float32; A the
        reference's sparse COO (CPU or CUDA, int8 or float) of shape (rows, R*N) — or a prebuilt RelGraph;
        A_idx the column subset of mini-batch mode.  `row_mask`/`relu` are optional fused extras used by RGCN."""
        dev = self._device()
        R = self.num_relations
        gI = gF = None
        wI = cI = wF = cF = None
        if self.input_layer:
            gI = graph_of(A, R, dev)
            wI, cI = self.weight_I, self.weight_I_comp
        Xd = None
        if not (self.input_layer and self.featureless):
            if X is None:
                raise ValueError("X is required unless the layer is a featureless input layer")
            Xd = X.to(dev).float()
            wF, cF = self.weight_F, self.weight_F_comp
            if A_idx is not None:
                row, col, val = slice_columns_device(A, A_idx, dev)
                gF = RelGraph.from_coo_arrays(row, col, val, A.shape[0], len(A_idx), R)
            else:
                gF = graph_of(A, R, dev)
        if row_mask is not None:
            row_mask = row_mask.to(dev).float()
        return _LayerFn.apply(Xd, wI, cI, wF, cF, self.b, row_mask, gI, gF, self.num_bases, relu, addend)
