"""Device-resident relational graph: the reference's stacked adjacency re-emitted as three sorted
edge orders (include/mrgcn_b200.h).

The reference hands its layer a torch sparse COO tensor `A` of shape (rows, R*N) built by
`scipy_sparse_to_pytorch_sparse` (/root/reference/mrgcn/data/utils.py:165-170, called with int8 from
mrgcn/data/batch.py:144-149) and lets `torch.mm` re-sort it on every call.  `RelGraph.from_coo`
consumes exactly that tensor once (values are taken from the tensor as handed over, so the int8
truncation of the reference is reproduced bit for bit) and `graph_of(A, R)` caches the result on the
tensor object, whose lifetime is that of the batch (batches are built once and reused every epoch,
mrgcn/tasks/node_classification.py:128-134).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nv

LONG_THRESH = 128      # rows / sources with more edges than this ("hubs") are handled by CTAs of their own
SLAB_ROWS = 32768      # minimum sources per slab of the relation-major order (a slab of feature rows stays L2-resident)
SLAB_GROUP_EDGES = 1024  # ... grown until a (slab, relation) group holds about this many edges on average
LONG_SEG = 512         # ... one CTA per segment of this many edges; partial sums combined in segment order
_I32 = torch.int32


def _chunk_size(E):
    # E3 chunks: enough of them to fill 148 SMs several times over, large enough to amortise staging
    ch = E // (148 * 24)
    return int(min(1024, max(128, (ch // 32) * 32)))


def chunk_worklist(grpptr, R, ch):
    """Work list of the relation-major kernels.  grpptr: edge range of every group g = slab*R + rel (E3 order).
    Every group is cut into chunks of at most `ch` edges (a chunk never crosses a group, hence holds one relation).
    Returns (chunk_rel[n], chunk_ptr[n+1], rel_chunk_ptr[R+1], rel_chunk_idx[n]); rel_chunk_idx lists the chunks of
    each relation in slab order, which fixes the order of the per-relation reductions."""
    grpptr = np.asarray(grpptr, dtype=np.int64)
    ngrp = len(grpptr) - 1
    cnt = np.diff(grpptr)
    nch = (cnt + ch - 1) // ch
    grp_chunk_ptr = np.zeros(ngrp + 1, dtype=np.int64)
    np.cumsum(nch, out=grp_chunk_ptr[1:])
    n_chunks = int(grp_chunk_ptr[-1])
    chunk_grp = np.repeat(np.arange(ngrp), nch)
    within = np.arange(n_chunks) - grp_chunk_ptr[chunk_grp]
    chunk_ptr = np.empty(n_chunks + 1, dtype=np.int64)
    chunk_ptr[:-1] = grpptr[chunk_grp] + within * ch
    chunk_ptr[-1] = grpptr[-1]
    chunk_rel = chunk_grp % R
    order = np.argsort(chunk_rel, kind="stable")
    rel_chunk_ptr = np.zeros(R + 1, dtype=np.int64)
    np.cumsum(np.bincount(chunk_rel, minlength=R), out=rel_chunk_ptr[1:])
    return chunk_rel, chunk_ptr, rel_chunk_ptr, order


def chunk_worklist_device(grpptr, R, ch):
    """`chunk_worklist` with torch ops on the tensor's device (no host round trip of the group pointers: mini-batch
    mode builds a graph per layer per batch).  Returns int64 tensors (chunk_rel, chunk_ptr, rel_chunk_ptr, order)."""
    grpptr = grpptr.long()
    dev = grpptr.device
    ngrp = grpptr.numel() - 1
    cnt = grpptr[1:] - grpptr[:-1]
    nch = (cnt + ch - 1) // ch
    grp_chunk_ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(nch, 0)])
    chunk_grp = torch.repeat_interleave(torch.arange(ngrp, device=dev), nch)
    n_chunks = chunk_grp.numel()
    within = torch.arange(n_chunks, device=dev) - grp_chunk_ptr[chunk_grp]
    chunk_ptr = torch.cat([grpptr[chunk_grp] + within * ch, grpptr[-1:]])
    chunk_rel = chunk_grp % R
    order = torch.argsort(chunk_rel, stable=True)
    rel_chunk_ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(torch.bincount(chunk_rel, minlength=R), 0)])
    return chunk_rel, chunk_ptr, rel_chunk_ptr, order


def hub_segments_device(deg, seg):
    """`hub_segments` with torch ops on the tensor's device."""
    d = deg.long()
    nseg = (d + seg - 1) // seg
    first = torch.cat([torch.zeros(1, dtype=torch.long, device=d.device), torch.cumsum(nseg, 0)])
    return torch.repeat_interleave(torch.arange(d.numel(), device=d.device), nseg), first


def hub_segments(deg, seg):
    """Hubs (rows / sources with `deg` entries each) cut into segments of `seg` entries: (seg_hub[n], seg_first[h+1])."""
    d = np.asarray(deg, dtype=np.int64)
    nseg = (d + seg - 1) // seg
    first = np.zeros(len(d) + 1, dtype=np.int64)
    np.cumsum(nseg, out=first[1:])
    return np.repeat(np.arange(len(d)), nseg), first


TAB_LT = 32          # edges per task of the table-term kernels (csrc/tab.cu)
TAB_TILE = 224       # a tile starts a new one every TAB_TILE edges of E2: at most TAB_TILE + TAB_LT - 1 edges per tile
TAB_PIECE = 32       # edges per (tile, relation) piece of the comp-gradient reduction
TAB_BLOCK = 128      # pieces per block of the two-stage sum of the piece records


def build_tab_plan(colptr, e2_rel, E, NS, R, long_thresh, lt=TAB_LT, tile=TAB_TILE, piece=TAB_PIECE, block=TAB_BLOCK):
    """Work plan of the table-term kernels (include/mrgcn_b200.h: mrgcn_tab_plan) from the source-major order E2.
    Pure torch, any device (the CPU tests check its invariants).  colptr: [NS+1], e2_rel: [>=E] integer tensors."""
    dev = colptr.device
    i32 = lambda t: t.to(torch.int32).contiguous()
    colptr = colptr.long()
    deg = colptr[1:] - colptr[:-1]
    ar = lambda n: torch.arange(n, device=dev)
    # tasks: up to `lt` consecutive edges of one source
    nt = (deg + lt - 1) // lt
    first = torch.cumsum(nt, 0) - nt
    n_tasks = int(nt.sum())
    task_src = torch.repeat_interleave(ar(NS), nt)
    task_lo = colptr[task_src] + lt * (ar(n_tasks) - first[task_src])
    # basis-gradient tasks: one per source that is not a hub, in node order (neighbouring lanes own neighbouring rows of
    # the gradient: their 8-byte stores fill whole sectors; ordering by degree was measured to turn them into
    # read-modify-write traffic, profiles/r02d_ncu_full_summary.txt)
    wsrc = torch.nonzero(deg <= long_thresh).flatten()
    task_len = torch.minimum(torch.full_like(task_lo, lt), colptr[task_src + 1] - task_lo)
    zero = torch.zeros_like(task_lo)
    tasks4 = torch.stack([task_src, task_lo, task_len, zero], 1)                  # one 16-byte descriptor per task
    wtasks4 = torch.stack([wsrc, colptr[wsrc], deg[wsrc], torch.zeros_like(wsrc)], 1)
    plan = dict(n_tasks=n_tasks, lt=lt, task_src=i32(task_src), task_lo=i32(task_lo), wsrc=i32(wsrc), n_wsrc=len(wsrc),
                tasks4=i32(tasks4), wtasks4=i32(wtasks4))
    if E == 0:
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        zr = torch.zeros(R + 1, dtype=torch.int32, device=dev)
        plan.update(n_tiles=0, n_pieces=0, n_blks=0, tile_slots=lt, tile_task_ptr=z, tile_e0=z, tperm=z, piece_ptr=z,
                    tile_piece_ptr=z, rel_piece_ptr=zr, rel_piece_idx=z, blk_ptr=z, rel_blk_ptr=zr)
        return plan
    # tiles: consecutive tasks; a new tile starts whenever a task starts in the next block of `tile` edges
    _, counts = torch.unique_consecutive(task_lo // tile, return_counts=True)
    n_tiles = len(counts)
    tile_task_ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(counts, 0)])
    tile_e0 = task_lo[tile_task_ptr[:-1]]
    tile_end = torch.cat([tile_e0[1:], torch.tensor([E], device=dev)])
    tile_slots = int((tile_end - tile_e0).max())
    tile_of_edge = torch.repeat_interleave(ar(n_tiles), tile_end - tile_e0)
    # tile-local relation order, cut into pieces of at most `piece` edges of one relation
    rel = e2_rel[:E].long()
    order = torch.argsort(tile_of_edge * R + rel, stable=True)          # E2 positions in (tile, relation, position) order
    ks = (tile_of_edge * R + rel)[order]
    idx = ar(E)
    newrun = torch.ones(E, dtype=torch.bool, device=dev)
    newrun[1:] = ks[1:] != ks[:-1]
    run_start = torch.cummax(torch.where(newrun, idx, torch.zeros_like(idx)), 0).values
    newpiece = ((idx - run_start) % piece) == 0
    piece_start = torch.nonzero(newpiece).flatten()
    n_pieces = len(piece_start)
    piece_ptr = torch.cat([piece_start, torch.tensor([E], device=dev)])
    tperm = order - tile_e0[tile_of_edge[order]]
    piece_tile = tile_of_edge[order][piece_start]
    tile_piece_ptr = torch.searchsorted(piece_tile, ar(n_tiles + 1))
    piece_rel = rel[order][piece_start]
    rel_piece_idx = torch.argsort(piece_rel, stable=True)
    rel_piece_ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(torch.bincount(piece_rel, minlength=R), 0)])
    # blocks of at most `block` consecutive entries of one relation's piece list (first stage of the record sum)
    per_rel = rel_piece_ptr[1:] - rel_piece_ptr[:-1]
    nblk = (per_rel + block - 1) // block
    rel_blk_ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(nblk, 0)])
    n_blks = int(rel_blk_ptr[-1])
    blk_rel = torch.repeat_interleave(ar(R), nblk)
    blk_lo = rel_piece_ptr[blk_rel] + block * (ar(n_blks) - rel_blk_ptr[blk_rel])
    blk_ptr = torch.cat([blk_lo, torch.tensor([n_pieces], device=dev)])
    plan.update(n_tiles=n_tiles, n_pieces=n_pieces, n_blks=n_blks, tile_slots=tile_slots, tile_task_ptr=i32(tile_task_ptr),
                tile_e0=i32(tile_e0), tperm=i32(tperm), piece_ptr=i32(piece_ptr), tile_piece_ptr=i32(tile_piece_ptr),
                rel_piece_ptr=i32(rel_piece_ptr), rel_piece_idx=i32(rel_piece_idx), blk_ptr=i32(blk_ptr), rel_blk_ptr=i32(rel_blk_ptr))
    return plan


class RelGraph:
    """E1/E2/E3 edge orders of one adjacency on one CUDA device."""

    def __init__(self, E, ND, NS, R, device):
        self.E, self.ND, self.NS, self.R, self.device = int(E), int(ND), int(NS), int(R), device
        e = self.E + 8          # slack: the TMA-engine kernels copy edge ranges rounded up to 16 bytes
        mk_i = lambda n: torch.empty(n, dtype=_I32, device=device)
        mk_f = lambda n: torch.empty(n, dtype=torch.float32, device=device)
        # slabs as small as L2 residency wants, but not so small that the (slab, relation) groups -- the unit of the
        # relation-major kernels -- degenerate (sparse shards of a partitioned graph)
        want = SLAB_ROWS
        if self.E > 0:
            while want < NS and (NS // want + 1) * R * SLAB_GROUP_EDGES > self.E:
                want *= 2
        self.slab_rows = int(want)
        self.n_slabs = max(1, -(-NS // self.slab_rows))
        self.rowptr, self.colptr, self.relptr = mk_i(ND + 1), mk_i(NS + 1), mk_i(self.n_slabs * R + 1)
        self.e1_src, self.e1_rel, self.e1_val = mk_i(e), mk_i(e), mk_f(e)
        self.e1_to_e2, self.e1_to_e3 = mk_i(e), mk_i(e)
        self.e2_src, self.e2_dst, self.e2_rel, self.e2_val = mk_i(e), mk_i(e), mk_i(e), mk_f(e)
        self.e2_to_e3 = mk_i(e)
        self.e3_src, self.e3_dst, self.e3_val, self.e3_to_e2 = mk_i(e), mk_i(e), mk_f(e), mk_i(e)
        self.long_rows = self.long_cols = None
        self.chunk_rel = self.chunk_ptr = self.rel_chunk_ptr = self.rel_chunk_idx = None
        self.n_chunks = 0
        self.c = nv.Graph()

    # ------------------------------------------------------------------------------------------
    def _fill_struct(self):
        c = self.c
        c.E, c.ND, c.NS, c.R = self.E, self.ND, self.NS, self.R
        c.slab_rows = self.slab_rows
        for name in ("rowptr", "e1_src", "e1_rel", "e1_val", "e1_to_e2", "e1_to_e3", "colptr", "e2_src", "e2_dst",
                     "e2_rel", "e2_val", "e2_to_e3", "relptr", "e3_src", "e3_dst", "e3_val", "e3_to_e2"):
            setattr(c, name, getattr(self, name).data_ptr())

    def _build_worklists(self, chunk=None):
        """Hub lists and the relation-chunk work list, computed on the device; the only host traffic is the handful of
        counts the launches need (one small device -> host copy per graph)."""
        dev = self.device
        deg_r = self.rowptr[1:] - self.rowptr[:-1]
        deg_c = self.colptr[1:] - self.colptr[:-1]
        self.long_rows = torch.nonzero(deg_r > LONG_THRESH).flatten().to(_I32)
        self.long_cols = torch.nonzero(deg_c > LONG_THRESH).flatten().to(_I32)
        one = torch.zeros(1, dtype=_I32, device=dev)

        def segments(deg, hubs):
            hub, first = hub_segments_device(deg[hubs.long()], LONG_SEG)
            return (hub.to(_I32) if hub.numel() else one), first.to(_I32), first[-1:]
        self.row_seg_hub, self.row_seg_first, nrs = segments(deg_r, self.long_rows)
        self.col_seg_hub, self.col_seg_first, ncs = segments(deg_c, self.long_cols)
        ch = chunk or _chunk_size(self.E)
        chunk_rel, chunk_ptr, rel_chunk_ptr, order = chunk_worklist_device(self.relptr, self.R, ch)
        n_chunks = chunk_rel.numel()          # a shape: known on the host without a copy
        self.n_row_segs, self.n_col_segs = (int(v) for v in torch.cat([nrs, ncs]).tolist())
        self.chunk_size = ch
        self.n_chunks = n_chunks
        self.chunk_rel = chunk_rel.to(_I32) if n_chunks else one
        self.chunk_ptr = chunk_ptr.to(_I32)
        self.rel_chunk_ptr = rel_chunk_ptr.to(_I32)
        self.rel_chunk_idx = order.to(_I32) if n_chunks else one
        c = self.c
        c.long_rows = self.long_rows.data_ptr() if len(self.long_rows) else None
        c.n_long_rows, c.long_row_thresh = len(self.long_rows), LONG_THRESH
        c.long_cols = self.long_cols.data_ptr() if len(self.long_cols) else None
        c.n_long_cols, c.long_col_thresh = len(self.long_cols), LONG_THRESH
        c.row_seg_hub, c.row_seg_first = self.row_seg_hub.data_ptr(), self.row_seg_first.data_ptr()
        c.col_seg_hub, c.col_seg_first = self.col_seg_hub.data_ptr(), self.col_seg_first.data_ptr()
        c.n_row_segs, c.n_col_segs, c.long_seg = self.n_row_segs, self.n_col_segs, LONG_SEG
        c.chunk_rel, c.chunk_ptr = self.chunk_rel.data_ptr(), self.chunk_ptr.data_ptr()
        c.rel_chunk_ptr, c.n_chunks = self.rel_chunk_ptr.data_ptr(), n_chunks
        c.rel_chunk_idx = self.rel_chunk_idx.data_ptr()
        # rows / columns by falling length: the warps of the one-pass narrow-layer kernels (csrc/narrow.cu) then hold rows
        # of (nearly) equal length instead of waiting for the longest of eight
        self.rows_by_deg = torch.argsort(deg_r, descending=True, stable=True).to(_I32) if self.ND else one
        self.cols_by_deg = torch.argsort(deg_c, descending=True, stable=True).to(_I32) if self.NS else one
        c.rows_by_deg, c.cols_by_deg = self.rows_by_deg.data_ptr(), self.cols_by_deg.data_ptr()

    # ------------------------------------------------------------------------------------------
    @classmethod
    def from_coo_arrays(cls, row, col, val, nrows, ncols, R, chunk=None):
        """row, col: int64 CUDA tensors; val: float32 CUDA tensor (already the values `A.float()` would give)."""
        for t, n in ((row, "row"), (col, "col"), (val, "val")):
            nv.require_cuda(t, n)
        if ncols % R:
            raise ValueError("adjacency has %d columns, not a multiple of num_relations=%d" % (ncols, R))
        row, col, val = row.contiguous(), col.contiguous(), val.contiguous().float()
        g = cls(row.numel(), nrows, ncols // R, R, row.device)
        g._fill_struct()
        with torch.cuda.device(row.device):
            nv.check(nv.lib().mrgcn_graph_build(nv.ptr(row), nv.ptr(col), nv.ptr(val), g.E, nrows, ncols, R,
                                                C.byref(g.c), nv.stream_ptr()), "graph_build")
            g._build_worklists(chunk)
        return g

    @classmethod
    def from_coo(cls, A, R, device=None, chunk=None):
        """A: torch sparse COO (int8 as the reference makes it, or float), CPU or CUDA; shape (rows, R*NS)."""
        if A.layout != torch.sparse_coo:
            raise TypeError("expected a torch sparse COO tensor (mrgcn/data/utils.py:165-170)")
        device = torch.device(device) if device is not None else (A.device if A.is_cuda else torch.device("cuda"))
        idx = A._indices().to(device, non_blocking=True)
        val = A._values().to(device, non_blocking=True).float()     # graph.py:75 `A.float()`: exact for int8
        return cls.from_coo_arrays(idx[0], idx[1], val, A.shape[0], A.shape[1], R, chunk)

    @classmethod
    def from_csr(cls, A, R, device="cuda", value_dtype=None, chunk=None):
        """The stacked adjacency as the reference stores it (scipy CSR from the tarball's A.npz: data / indices / indptr,
        /root/reference/mrgcn/data/io/tarball.py:151-157) straight to the device edge orders, without the detour through
        `.nonzero()` and an int64 COO tensor on the host.  value_dtype=torch.int8 reproduces the truncation of
        FullBatch.as_tensors_ (mrgcn/data/batch.py:148-149: 1/deg -> 0 for deg >= 2); default keeps the float values."""
        device = torch.device(device)
        indptr = torch.from_numpy(np.asarray(A.indptr, dtype=np.int64)).to(device)
        col = torch.from_numpy(np.asarray(A.indices, dtype=np.int64)).to(device)
        val = torch.from_numpy(np.asarray(A.data, dtype=np.float32)).to(device)
        if value_dtype is not None:
            val = val.to(value_dtype).float()
        row = torch.repeat_interleave(torch.arange(A.shape[0], device=device), indptr[1:] - indptr[:-1])
        return cls.from_coo_arrays(row, col, val, A.shape[0], A.shape[1], R, chunk)

    @classmethod
    def from_triples(cls, triples, num_nodes, num_props, include_inverse=True, device="cuda", chunk=None):
        """Integer triples (s,p,o) -> normalised stacked adjacency built on the GPU, bit-exact with
        mrgcn/encodings/graph_structure.py:70-108,162-169 + the float32 cast of tarball.py:151-157."""
        device = torch.device(device)
        tr = torch.as_tensor(np.ascontiguousarray(triples, dtype=np.int32)).to(device)
        T = tr.shape[0]
        R = (2 * num_props if include_inverse else num_props) + 1
        n = (2 * T if include_inverse else T) + num_nodes
        row = torch.empty(n, dtype=torch.int64, device=device)
        col = torch.empty(n, dtype=torch.int64, device=device)
        val = torch.empty(n, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            nv.check(nv.lib().mrgcn_adjacency_from_triples(nv.ptr(tr), T, num_nodes, num_props, int(include_inverse),
                                                           nv.ptr(row), nv.ptr(col), nv.ptr(val), nv.stream_ptr()),
                     "adjacency_from_triples")
        g = cls.from_coo_arrays(row, col, val, num_nodes, R * num_nodes, R, chunk)
        g.coo = (row, col, val)
        return g

    def _row_edges(self, rows):
        """E1 positions of the entries of `rows` (device int64), and the local row of every entry."""
        rows = rows.to(self.device).long()
        lo = self.rowptr[rows].long()
        deg = self.rowptr[rows + 1].long() - lo
        total_first = torch.cumsum(deg, 0) - deg
        local = torch.repeat_interleave(torch.arange(rows.numel(), device=self.device), deg)
        pos = torch.arange(local.numel(), device=self.device) - total_first[local] + lo[local]
        return pos, local

    def row_slice(self, rows, value_dtype=None):
        """`A[rows]` (mrgcn/data/batch.py:190) as a device sparse COO of shape (len(rows), R*NS), without leaving the GPU.
        value_dtype=torch.int8 reproduces the truncation of A_Batch.as_tensors_ (batch.py:223-226)."""
        pos, local = self._row_edges(rows)
        col = self.e1_rel[pos].long() * self.NS + self.e1_src[pos].long()
        val = self.e1_val[pos]
        if value_dtype is not None:
            val = val.to(value_dtype)
        return torch.sparse_coo_tensor(torch.stack([local, col]), val, (int(rows.numel()), self.R * self.NS))

    def neighbours(self, rows):
        """`getNeighboursSparse` (batch.py:228-243) on the device: the sorted set of source nodes of `rows`, any relation."""
        pos, _ = self._row_edges(rows)
        return torch.unique(self.e1_src[pos].long())

    def tab_plan(self):
        """Work plan of the table-term kernels over this graph's E2 order (built on first use, then cached)."""
        if getattr(self, "_tab", None) is None:
            with torch.cuda.device(self.device):
                d = build_tab_plan(self.colptr, self.e2_rel, self.E, self.NS, self.R, LONG_THRESH)
            c = nv.TabPlan()
            for k in ("n_tasks", "n_wsrc", "n_tiles", "n_pieces", "tile_slots", "lt", "n_blks"):
                setattr(c, k, int(d[k]))
            for k in ("task_src", "task_lo", "tasks4", "wsrc", "wtasks4", "tile_task_ptr", "tile_e0", "tperm", "piece_ptr", "tile_piece_ptr",
                      "rel_piece_ptr", "rel_piece_idx", "blk_ptr", "rel_blk_ptr"):
                if d[k].numel() == 0:
                    d[k] = torch.zeros(1, dtype=_I32, device=self.device)
                setattr(c, k, d[k].data_ptr())
            self._tab = (c, d)
        return self._tab[0]

    def to_coo(self, dtype=torch.float32):
        """Back to a (coalesced-order) torch sparse COO on the device: E1 order, column = rel*NS + src."""
        deg = (self.rowptr[1:] - self.rowptr[:-1]).long()
        row = torch.repeat_interleave(torch.arange(self.ND, device=self.device), deg)
        col = self.e1_rel[:self.E].long() * self.NS + self.e1_src[:self.E].long()
        return torch.sparse_coo_tensor(torch.stack([row, col]), self.e1_val[:self.E].to(dtype), (self.ND, self.R * self.NS))

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in vars(self).values() if isinstance(t, torch.Tensor))


def graph_of(A, R, device=None):
    """RelGraph of a reference-style sparse COO tensor, cached on the tensor object itself."""
    if isinstance(A, RelGraph):
        return A
    cache = getattr(A, "_mrgcn_b200_graph", None)
    dev = torch.device(device) if device is not None else None
    if cache is not None and cache[0] == (R, A._version, A._nnz()) and (dev is None or cache[1].device == dev or
                                                                         (dev.index is None and cache[1].device.type == dev.type)):
        return cache[1]
    g = RelGraph.from_coo(A, R, device)
    try:
        A._mrgcn_b200_graph = ((R, A._version, A._nnz()), g)
    except Exception:   # tensors that refuse attributes: rebuild per call
        pass
    return g
