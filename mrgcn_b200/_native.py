"""ctypes binding of libmrgcn_b200.so (C ABI declared in include/mrgcn_b200.h).

There is no CPU fallback: importing this module without the built library, or calling into it
without a CUDA device, raises.  Build with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C mrgcn_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libmrgcn_b200.so")

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)


class Graph(C.Structure):
    """struct mrgcn_graph (include/mrgcn_b200.h)."""
    _fields_ = [
        ("E", C.c_int64), ("ND", C.c_int32), ("NS", C.c_int32), ("R", C.c_int32), ("_pad", C.c_int32),
        ("rowptr", C.c_void_p), ("e1_src", C.c_void_p), ("e1_rel", C.c_void_p), ("e1_val", C.c_void_p),
        ("e1_to_e2", C.c_void_p), ("e1_to_e3", C.c_void_p),
        ("colptr", C.c_void_p), ("e2_src", C.c_void_p), ("e2_dst", C.c_void_p), ("e2_rel", C.c_void_p),
        ("e2_val", C.c_void_p), ("e2_to_e3", C.c_void_p),
        ("relptr", C.c_void_p), ("e3_src", C.c_void_p), ("e3_dst", C.c_void_p), ("e3_val", C.c_void_p),
        ("e3_to_e2", C.c_void_p),
        ("long_rows", C.c_void_p), ("n_long_rows", C.c_int32), ("long_row_thresh", C.c_int32),
        ("long_cols", C.c_void_p), ("n_long_cols", C.c_int32), ("long_col_thresh", C.c_int32),
        ("row_seg_hub", C.c_void_p), ("row_seg_first", C.c_void_p), ("col_seg_hub", C.c_void_p),
        ("col_seg_first", C.c_void_p), ("n_row_segs", C.c_int32), ("n_col_segs", C.c_int32), ("long_seg", C.c_int32),
        ("_pad3", C.c_int32),
        ("chunk_rel", C.c_void_p), ("chunk_ptr", C.c_void_p), ("rel_chunk_ptr", C.c_void_p),
        ("rel_chunk_idx", C.c_void_p), ("n_chunks", C.c_int32), ("slab_rows", C.c_int32),
        ("rows_by_deg", C.c_void_p), ("cols_by_deg", C.c_void_p),
    ]


class TabPlan(C.Structure):
    """struct mrgcn_tab_plan."""
    _fields_ = [
        ("n_tasks", C.c_int32), ("n_wsrc", C.c_int32), ("n_tiles", C.c_int32), ("n_pieces", C.c_int32),
        ("tile_slots", C.c_int32), ("lt", C.c_int32), ("_pad0", C.c_int32), ("_pad1", C.c_int32),
        ("task_src", C.c_void_p), ("task_lo", C.c_void_p), ("tasks4", C.c_void_p), ("wsrc", C.c_void_p),
        ("wtasks4", C.c_void_p), ("tile_task_ptr", C.c_void_p),
        ("tile_e0", C.c_void_p), ("tperm", C.c_void_p), ("piece_ptr", C.c_void_p), ("tile_piece_ptr", C.c_void_p),
        ("rel_piece_ptr", C.c_void_p), ("rel_piece_idx", C.c_void_p), ("blk_ptr", C.c_void_p), ("rel_blk_ptr", C.c_void_p),
        ("n_blks", C.c_int32), ("_pad2", C.c_int32),
    ]


class LayerArgs(C.Structure):
    """struct mrgcn_layer_args."""
    _fields_ = [
        ("gI", C.POINTER(Graph)), ("gF", C.POINTER(Graph)),
        ("in_dim", C.c_int32), ("out_dim", C.c_int32), ("B", C.c_int32), ("relu", C.c_int32),
        ("weight_I", C.c_void_p), ("comp_I", C.c_void_p), ("X", C.c_void_p), ("weight_F", C.c_void_p),
        ("comp_F", C.c_void_p), ("bias", C.c_void_p), ("row_mask", C.c_void_p), ("addend", C.c_void_p),
        ("wmix", C.c_void_p), ("msg_I", C.c_void_p), ("msg_F", C.c_void_p), ("hub_ws", C.c_void_p),
        ("out", C.c_void_p),
        ("plan", C.POINTER(TabPlan)), ("proj", C.c_void_p), ("vt_ws", C.c_void_p), ("xpad_ws", C.c_void_p),
        ("x_stride", C.c_int32), ("_pad", C.c_int32),
    ]


class LayerBwdArgs(C.Structure):
    """struct mrgcn_layer_bwd_args."""
    _fields_ = [
        ("f", LayerArgs), ("gout", C.c_void_p),
        ("g_weight_I", C.c_void_p), ("g_comp_I", C.c_void_p), ("g_weight_F", C.c_void_p),
        ("g_comp_F", C.c_void_p), ("g_bias", C.c_void_p), ("g_X", C.c_void_p),
        ("gact", C.c_void_p), ("cbuf", C.c_void_p), ("part", C.c_void_p), ("g_wmix", C.c_void_p),
        ("colsum_ws", C.c_void_p), ("wt_ws", C.c_void_p), ("msgx_ws", C.c_void_p),
        ("phases", C.c_int32), ("_pad", C.c_int32),
    ]


BWD_ACT, BWD_IDENT, BWD_FEATW, BWD_GX = 1, 2, 4, 8


# name -> (restype, argtypes); every symbol include/mrgcn_b200.h declares
SYMBOLS = {
    "mrgcn_version": (C.c_int, []),
    "mrgcn_last_error_string": (C.c_char_p, []),
    "mrgcn_launch_count": (C.c_int64, []),
    "mrgcn_profile_enable": (None, [C.c_int]),
    "mrgcn_profile_dump": (C.c_int64, [C.c_char_p, C.c_int64]),
    "mrgcn_graph_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                    C.POINTER(Graph), C.c_void_p]),
    "mrgcn_adjacency_from_triples": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "mrgcn_msg_stride": (C.c_int32, [C.c_int32]),
    "mrgcn_tab_mode": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "mrgcn_set_tab_mask": (None, [C.c_int32]),
    "mrgcn_set_ident_fused": (None, [C.c_int32]),
    "mrgcn_feat_proj_supported": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "mrgcn_feat_proj": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "mrgcn_upload_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "mrgcn_rgcn_layer_fwd": (C.c_int, [C.POINTER(LayerArgs), C.c_void_p]),
    "mrgcn_rgcn_layer_bwd": (C.c_int, [C.POINTER(LayerBwdArgs), C.c_void_p]),
    "mrgcn_sqnorm_ws_elems": (C.c_int64, []),
    "mrgcn_grad_sqnorm": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mrgcn_adam_clip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float,
                                  C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, C.c_void_p]),
    "mrgcn_scatter_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p]),
    "mrgcn_scatter_rows_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "mrgcn_distmult_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
    "mrgcn_distmult_bwd_ws_elems": (C.c_int64, [C.c_int64]),
    "mrgcn_distmult_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mrgcn_narrow_supported": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "mrgcn_distmult_rank_ws_elems": (C.c_int64, [C.c_int64]),
    "mrgcn_distmult_rank": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("mrgcn_b200: %s is missing - build it (make -C mrgcn_b200/csrc); there is no "
                               "CPU fallback" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().mrgcn_last_error_string().decode("utf-8", "replace")
        raise RuntimeError("mrgcn_b200 %s failed (code %d): %s" % (what, rc, msg))


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("mrgcn_b200: %s must live on a CUDA device (got %s); there is no CPU path" % (name, t.device))


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().mrgcn_launch_count())


def profile_enable(on=True):
    lib().mrgcn_profile_enable(int(bool(on)))


def profile_dump():
    """{kernel name: (launches, total_ms)} since the last dump (synchronises the device)."""
    buf = C.create_string_buffer(1 << 16)
    lib().mrgcn_profile_dump(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.rsplit(" ", 2)
        out[name] = (int(n), float(ms))
    return out
