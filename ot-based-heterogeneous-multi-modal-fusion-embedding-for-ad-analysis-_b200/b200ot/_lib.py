"""ctypes binding of libb200ot.so (the C ABI declared in include/b200ot.h).

There is no CPU fallback: if the shared object is missing the import of the
compute entry points fails loudly with instructions to build it.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200OT_LIB lets a tuning run A/B two builds of the same ABI in one process launch each
LIB_PATH = os.environ.get("B200OT_LIB") or os.path.join(HERE, "libb200ot.so")

# status codes / enums (include/b200ot.h)
OK, E_INVALID, E_WORKSPACE, E_LAUNCH, E_UNSUPPORTED, E_NUMERIC = 0, -1, -2, -3, -4, -5
NORM_L2, NORM_L2SQ, NORM_L1 = 0, 1, 2
PATH_AUTO, PATH_FUSED, PATH_ROBUST = 0, 1, 2
COST_SQEUCLIDEAN, COST_COSINE = 0, 1
OT_COST_DOUBLES = 1 + 148 * 8 + 1  # B200OT_OT_COST_DOUBLES
NORMS = {"l2": NORM_L2, "l2sq": NORM_L2SQ, "l1": NORM_L1}
PATHS = {"auto": PATH_AUTO, "fused": PATH_FUSED, "robust": PATH_ROBUST}
TERMS_F16_3, TERMS_F16_4 = 19, 20  # b200ot.h: two fp16 parts per operand, 3 / 4 products


def split_terms(terms):
    """(code for the C ABI, tensor products per element, fp16?) of a `terms` argument: 1 / 3 / 6 = bf16 parts,
    "f16" or 19 = two fp16 parts and 3 products, "f16x4" or 20 = the same plus x2.y2."""
    table = {1: (1, 1, False), 3: (3, 3, False), 6: (6, 6, False), "f16": (19, 3, True), 19: (19, 3, True),
             "f16x4": (20, 4, True), 20: (20, 4, True)}
    if terms not in table:
        raise ValueError(f"terms must be one of {sorted(map(str, table))}, got {terms!r}")
    return table[terms]


COSTS = {"sqeuclidean": COST_SQEUCLIDEAN, "cosine": COST_COSINE}


class Params(C.Structure):
    _fields_ = [("eps", C.c_float), ("max_iter", C.c_int), ("tol", C.c_float),
                ("check_every", C.c_int), ("check_phase", C.c_int), ("err_norm", C.c_int),
                ("stop_inclusive", C.c_int), ("path", C.c_int), ("floor_patience", C.c_int)]


class Result(C.Structure):
    _fields_ = [("n_iter", C.c_int), ("converged", C.c_int), ("status", C.c_int),
                ("n_err", C.c_int), ("err", C.c_float), ("reserved", C.c_float * 3)]


class B200OTError(RuntimeError):
    pass


_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/b200ot.h one to one
SIGNATURES = {
    "b200ot_version": (_i, []),
    "b200ot_strerror": (C.c_char_p, [_i]),
    "b200ot_last_cuda_error": (C.c_char_p, []),
    "b200ot_cost_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200ot_cost": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _p, _i, _p, _sz, _i, _p]),
    "b200ot_cost_parts_bytes": (_sz, [_i, _i, _i]),
    "b200ot_cost_split": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "b200ot_cost_gemm": (_i, [_p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _p, _i, _p]),
    "b200ot_cost_simt": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _p, _i, _p, _p]),
    "b200ot_fot_cost": (_i, [_p, _i, _p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p, _i, _p, _p]),
    "b200ot_fot_cost_tc_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "b200ot_fot_cost_tc": (_i, [_p, _i, _p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p, _i, _p, _sz, _p]),
    "b200ot_matrix_max": (_i, [_p, _i, _i, _i, _p, _p]),
    "b200ot_matrix_scale_by_inv": (_i, [_p, _i, _i, _i, _p, _p]),
    "b200ot_sinkhorn_workspace_bytes": (_sz, [_i, _i]),
    "b200ot_sinkhorn_setup": (_i, [_i, _i, _p, _p, _p, _p, C.POINTER(Params), _p, _sz, _p]),
    "b200ot_sinkhorn_init": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, C.POINTER(Params), _p, _sz, _p]),
    "b200ot_sinkhorn_enqueue": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "b200ot_sinkhorn_snapshot": (_i, [_i, _i, _p, _p]),
    "b200ot_sinkhorn_rewind": (_i, [_i, _i, _p, _p]),
    "b200ot_sinkhorn_peek": (_i, [_p, _p, _p]),
    "b200ot_sinkhorn_counter": (C.c_longlong, [_i]),
    "b200ot_sinkhorn_describe": (_i, [_i, _i, C.c_char_p, _i]),
    "b200ot_sinkhorn_finish": (_i, [_i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "b200ot_sinkhorn_solve": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, C.POINTER(Params), _p, _sz, _p, _p,
                                   C.POINTER(Result), _p, _i, _p]),
    "b200ot_sinkhorn_shard_prologue": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "b200ot_sinkhorn_shard_sweep": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "b200ot_sinkhorn_shard_finalize": (_i, [_i, _i, _p, _p, _i, _p]),
    "b200ot_sinkhorn_panel_prologue": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "b200ot_sinkhorn_panel_sweep": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "b200ot_nccl_unique_id": (_i, [_p]),
    "b200ot_nccl_init": (_i, [_p, _i, _i, C.POINTER(C.c_void_p)]),
    "b200ot_nccl_destroy": (_i, [_p]),
    "b200ot_sinkhorn_shard_start": (_i, [_p, _i, _i, _i, _p, _p, _p, _p]),
    "b200ot_sinkhorn_shard_run": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "b200ot_peer_alloc": (_i, [_sz, C.POINTER(C.c_void_p), _p]),
    "b200ot_peer_open": (_i, [_p, C.POINTER(C.c_void_p)]),
    "b200ot_peer_close": (_i, [_p]),
    "b200ot_peer_free": (_i, [_p]),
    "b200ot_peer_exchange_bytes": (_sz, [_i, _i]),
    "b200ot_sinkhorn_shard_push": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _i, C.c_uint, _i, _p]),
    "b200ot_sinkhorn_shard_finalize_peer": (_i, [_i, _i, _p, _p, _i, C.c_uint, _i, _p]),
    "b200ot_sinkhorn_shard_run_peer": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _i, _i, C.c_uint, _p]),
    "b200ot_sinkhorn_batched": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, C.POINTER(Params), _p, _p, _p,
                                     _p, _p, _p]),
    "b200ot_plan": (_i, [_p, _i, _i, _i, _p, _p, _f, _p, _i, _p]),
    "b200ot_ot_cost": (_i, [_p, _i, _i, _i, _p, _p, _f, _p, _p]),
    "b200ot_plan_guard_rownorm": (_i, [_p, _i, _i, _i, _p, _p, _f, _p, _i, _p, _i, _p]),
    "b200ot_apply_plan": (_i, [_p, _i, _i, _i, _p, _p, _f, _p, _i, _i, _i, _p, _i, _p]),
    "b200ot_apply_plan_t": (_i, [_p, _i, _i, _i, _p, _p, _f, _p, _i, _i, _i, _p, _i, _p]),
    "b200ot_apply_plan_tc_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "b200ot_apply_plan_tc": (_i, [_p, _i, _i, _i, _p, _p, _f, _p, _i, _i, _i, _i, _p, _i, _p, _p, _sz, _p]),
    "b200ot_apply_plan_tc_weighted": (_i, [_p, _i, _i, _i, _p, _p, _f, _f, _f, _p, _p, _p, _i, _i, _i, _p, _i, _p, _p,
                                           _sz, _p]),
    "b200ot_envelope_bwd_weighted": (_i, [_p, _i, _i, _i, _p, _p, _f, _f, _f, _p, _p, _p, _i, _p, _i, _i, _f, _p, _i,
                                          _p, _i, _p, _sz, _p]),
    "b200ot_envelope_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "b200ot_envelope_bwd": (_i, [_p, _i, _i, _i, _p, _p, _f, _p, _i, _p, _i, _i, _f, _p, _i, _p, _i, _p, _p, _p, _sz, _p]),
    "b200ot_cosine_loss": (_i, [_p, _i, _p, _i, _i, _i, _p, _p]),
    "b200ot_foscttm": (_i, [_p, _i, _i, _p, _p]),
    "b200ot_egw_batched": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _f, _i, _i, _f, _p, _p, _p, _p]),
    "b200ot_token_attention_fwd": (_i, [_p, _i, _i, _i, _i, _p, _f, _p, _p, _p]),
    "b200ot_token_attention_bwd": (_i, [_p, _p, _p, _f, _p, _i, _i, _i, _i, _p, _p]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200OTError(
            f"{LIB_PATH} is missing. b200ot has no CPU or PyTorch fallback: build the CUDA "
            "library first (python -c 'import __graft_entry__ as g; g.build()' or "
            "python -m b200ot.build).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str = ""):
    if code == OK:
        return
    lib = load()
    msg = lib.b200ot_strerror(code).decode()
    if code == E_LAUNCH:
        msg += ": " + lib.b200ot_last_cuda_error().decode()
    raise B200OTError(f"{what or 'b200ot'} failed ({code}): {msg}")
