"""ctypes binding of libgode.so (the C ABI in include/gode.h).  No torch types cross this boundary:
device pointers are passed as integers, sizes as int64, the stream as the raw cudaStream_t handle."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libgode.so")

MAX_STAGES = 8
MAX_PEERS, PEER_HANDLE_BYTES, PEER_HEADER_BYTES, PEER_TIMEOUT = 16, 64, 4096, 1
HEAVY_ROW = 256
HEAVY_CHUNK = 256
PREC_FP32, PREC_TF32 = 0, 1
PROF_AGG_FWD, PROF_AGG_T, PROF_TRANSFORM, PROF_VJP_DENSE, PROF_OTHER = range(5)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libgode.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "from the repository root; there is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

vp, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t


class PushRoute(C.Structure):
    _fields_ = [("ptr", vp), ("ent", vp), ("base", vp * MAX_PEERS)]


class RkSecond(C.Structure):
    _fields_ = [("coef", f32 * MAX_STAGES), ("coef_self", f32), ("out", vp)]


class SpmmEpilogue(C.Structure):
    _fields_ = [("bias", vp), ("relu", i32), ("residual", vp), ("y0", vp), ("kprev", vp * MAX_STAGES),
                ("coef", f32 * MAX_STAGES), ("n_prev", i32), ("coef_self", f32), ("ynext", vp),
                ("mask_src", vp), ("mask_scale", f32), ("gp_out", vp), ("acc_in", vp), ("push", PushRoute),
                ("push_y", PushRoute), ("second", RkSecond), ("gp_row_scale", vp)]


class Csr(C.Structure):
    _fields_ = [("n_rows", i64), ("n_cols", i64), ("rowptr", vp), ("colidx", vp), ("vals", vp), ("row_vals", vp),
                ("heavy_rows", vp), ("heavy_chunk_ptr", vp), ("n_heavy", i32), ("n_chunks", i32), ("tile_sched", vp),
                ("n_tile_sched", i32)]


GAT_CHUNK = 64


class GatHeavy(C.Structure):
    _fields_ = [("n_heavy", i32), ("n_chunks", i32), ("nodes", vp), ("cptr", vp), ("chunk_node", vp), ("chunk_e0", vp)]


class GatGraph(C.Structure):
    _fields_ = [("n_nodes", i64), ("n_edges", i64), ("tptr", vp), ("t_src", vp), ("t_tgt", vp), ("sptr", vp),
                ("s_tgt", vp), ("s_pos", vp), ("t_heavy", GatHeavy), ("s_heavy", GatHeavy)]


class GcnOdeFunc(C.Structure):
    _fields_ = [("A", Csr), ("At", Csr), ("d", i32), ("groups", i32), ("gn_eps", f32), ("precision", i32),
                ("W", vp), ("b", vp), ("gamma", vp), ("beta", vp), ("gather_row_offset", i64), ("partial_in", vp),
                ("push_S", PushRoute), ("push_gP", PushRoute), ("push_y", PushRoute), ("second", RkSecond), ("gp_row_scale", vp)]


class PeerGroup(C.Structure):
    _fields_ = [("world", i32), ("rank", i32), ("base", vp * MAX_PEERS)]


_PROTOS = {
    "gode_version": (C.c_int, []),
    "gode_last_error": (C.c_char_p, []),
    "gode_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "gode_launch_count": (C.c_ulonglong, []),
    "gode_reserve_sms": (C.c_int, [C.c_int]),
    "gode_profile_enable": (C.c_int, [C.c_int]),
    "gode_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "gode_csr_from_coo_workspace_bytes": (sz, [i64, i64]),
    "gode_csr_from_coo": (C.c_int, [i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "gode_csr_transpose_workspace_bytes": (sz, [i64, i64, i64]),
    "gode_csr_transpose": (C.c_int, [i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "gode_csr_heavy_rows": (C.c_int, [i64, vp, vp, vp, vp, vp]),
    "gode_csr_row_values": (C.c_int, [i64, vp, vp, vp, vp, vp]),
    "gode_spmm_workspace_bytes": (sz, [C.POINTER(Csr), i32]),
    "gode_spmm_csr_f32": (C.c_int, [C.POINTER(Csr), vp, i64, i32, vp, i64, C.POINTER(SpmmEpilogue), vp, sz, vp]),
    "gode_gemm_f32": (C.c_int, [i32, i32, i64, i64, i64, f32, vp, i64, vp, i64, f32, vp, i64, i32, i32, vp, sz, vp]),
    "gode_linear_f32": (C.c_int, [i32, i64, i64, i64, vp, i64, vp, i64, vp, i32, vp, i64, vp]),
    "gode_gemm_tc_f32": (C.c_int, [i64, i64, i64, vp, i64, vp, i64, vp, i32, vp, i64, i32, vp]),
    "gode_lsm_nll_workspace_bytes": (sz, [i64]),
    "gode_lsm_nll_fwd": (C.c_int, [i64, i32, vp, i64, vp, vp, i64, vp, i64, vp, vp, sz, vp]),
    "gode_lsm_nll_bwd": (C.c_int, [i64, i32, vp, i64, vp, vp, i64, vp, vp, i64, vp]),
    "gode_adam_step": (C.c_int, [i64, vp, vp, vp, vp, f32, f32, f32, f32, f32, vp, vp]),
    "gode_gemm_tc_splitk_f32": (C.c_int, [i64, i64, i64, vp, i64, vp, i64, vp, i64, i32, i32, vp]),
    "gode_groupnorm_fwd": (C.c_int, [i64, i32, i32, f32, vp, i64, vp, vp, vp, i64, vp]),
    "gode_groupnorm_bwd": (C.c_int, [i64, i32, i32, f32, vp, i64, vp, vp, i64, vp, i64, vp, vp, vp, sz, vp]),
    "gode_qc_collate": (C.c_int, [i32, vp, vp, vp, vp, vp, i32, i64, i64, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]),
    "gode_gn_linear_supported": (C.c_int, [i32, i32, i32]),
    "gode_gn_linear_f32": (C.c_int, [i64, i32, i32, f32, vp, vp, vp, vp, i64, f32, i32, vp, i64, i32, vp]),
    "gode_gn_wgrad_workspace_bytes": (sz, [i32]),
    "gode_gn_wgrad_f32": (C.c_int, [i64, i32, i32, f32, vp, vp, i64, i32, vp, i64, vp, vp, sz, i32, vp]),
    "gode_colreduce_workspace_bytes": (sz, [i32]),
    "gode_colsum_f32": (C.c_int, [i64, i32, vp, i64, vp, vp, sz, vp]),
    "gode_relu_bwd": (C.c_int, [i64, vp, vp, vp, vp]),
    "gode_rk_combine": (C.c_int, [i64, vp, C.POINTER(vp), C.POINTER(f32), i32, vp, vp]),
    "gode_rk_error_sumsq": (C.c_int, [i64, vp, vp, C.POINTER(vp), C.POINTER(f32), i32, f32, f32, vp, vp, sz, vp]),
    "gode_gcn_workspace_bytes": (sz, [C.POINTER(GcnOdeFunc)]),
    "gode_gcn_push_fusable": (C.c_int, [C.POINTER(GcnOdeFunc)]),
    "gode_gcn_transform": (C.c_int, [C.POINTER(GcnOdeFunc), vp, f32, vp, vp, sz, vp]),
    "gode_gcn_transform_rows": (C.c_int, [C.POINTER(GcnOdeFunc), vp, f32, vp, i64, i64, vp, sz, vp]),
    "gode_gcn_stage_fwd": (C.c_int, [C.POINTER(GcnOdeFunc), vp, vp, vp, C.POINTER(vp), C.POINTER(f32), i32, f32, vp,
                                     f32, vp, vp, sz, vp]),
    "gode_gcn_stage_fwd_rows": (C.c_int, [C.POINTER(GcnOdeFunc), vp, vp, vp, C.POINTER(vp), C.POINTER(f32), i32, f32, vp,
                                          i64, i64, vp, sz, vp]),
    "gode_gcn_stage_vjp": (C.c_int, [C.POINTER(GcnOdeFunc), vp, f32, vp, vp, f32, vp, vp, vp, vp, sz, vp]),
    "gode_gcn_vjp_phase1": (C.c_int, [C.POINTER(GcnOdeFunc), vp, vp, f32, vp, vp, vp, C.POINTER(vp), C.POINTER(f32), i32,
                                      f32, vp, vp, sz, vp]),
    "gode_gcn_vjp_phase2": (C.c_int, [C.POINTER(GcnOdeFunc), vp, f32, vp, vp, vp, vp, sz, vp]),
    "gode_gcn_vjp_phase2_rk": (C.c_int, [C.POINTER(GcnOdeFunc), vp, f32, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(f32), i32,
                                         f32, vp, vp, sz, vp]),
    "gode_segment_attend_fwd": (C.c_int, [i32, i32, vp, vp, i64, vp, i64, vp, vp, i64, vp]),
    "gode_segment_attend_bwd": (C.c_int, [i32, i32, vp, vp, i64, vp, i64, vp, vp, i64, vp, i64, vp, i64, vp]),
    "gode_gather_rows": (C.c_int, [i64, vp, i32, vp, i64, vp, i64, vp]),
    "gode_peer_alloc": (C.c_int, [sz, C.POINTER(vp)]),
    "gode_peer_free": (C.c_int, [vp]),
    "gode_peer_export": (C.c_int, [vp, vp]),
    "gode_peer_open": (C.c_int, [vp, C.POINTER(vp)]),
    "gode_peer_close": (C.c_int, [vp]),
    "gode_halo_push": (C.c_int, [C.POINTER(PeerGroup), C.c_uint32, vp, vp, vp, i64, i32, vp, i64, i64, i32, vp]),
    "gode_halo_push_part": (C.c_int, [C.POINTER(PeerGroup), C.c_uint32, vp, vp, vp, vp, i64, i32, vp, i64, i64, i32, i32, vp]),
    "gode_peer_wait": (C.c_int, [C.POINTER(PeerGroup), C.c_uint32, C.c_uint64, vp]),
    "gode_peer_status": (C.c_int, [C.POINTER(PeerGroup), C.POINTER(i32), vp]),
    "gode_edge_matvec": (C.c_int, [i64, i32, vp, vp, vp, i64, vp, vp]),
    "gode_edge_matvec_bwd": (C.c_int, [i64, i32, vp, vp, vp, vp, i64, vp, i64, vp, vp, vp]),
    "gode_gat_fwd_workspace_bytes": (sz, [C.POINTER(GatGraph), i32, i32]),
    "gode_gat_fwd": (C.c_int, [C.POINTER(GatGraph), i32, i32, vp, i64, f32, vp, i64, vp, vp, vp, vp, sz, vp]),
    "gode_gat_bwd_workspace_bytes": (sz, [C.POINTER(GatGraph), i32, i32]),
    "gode_gat_bwd": (C.c_int, [C.POINTER(GatGraph), i32, i32, vp, i64, vp, i64, vp, vp, vp, i64, vp, vp, sz, vp]),
}

EXPORTS = tuple(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)  # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


class GodeError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib.gode_last_error()
        raise GodeError("%s failed (code %d): %s" % (what or "libgode call", rc, msg.decode() if msg else "?"))


def ptr_array(ptrs):
    arr = (vp * MAX_STAGES)()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def f32_array(vals):
    arr = (f32 * MAX_STAGES)()
    for i, v in enumerate(vals):
        arr[i] = v
    return arr
