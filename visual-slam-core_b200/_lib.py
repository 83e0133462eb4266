"""ctypes binding of include/b200ba.h.  No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import _build

BA_MAX_ITER = 64
BA_NCCL_ID_BYTES = 128
BA_OK = 0
BA_ERR_INVALID, BA_ERR_CUDA, BA_ERR_TOO_LARGE, BA_ERR_NCCL, BA_ERR_NO_DEVICE = -1, -2, -3, -4, -5
BA_SOLVER_AUTO, BA_SOLVER_CHOLESKY, BA_SOLVER_PCG, BA_SOLVER_BAND, BA_SOLVER_BAND_SEQ = 0, 1, 2, 3, 4
BA_SCHUR_AUTO, BA_SCHUR_SCATTER, BA_SCHUR_TILED, BA_SCHUR_STREAM, BA_SCHUR_MMA, BA_SCHUR_PANEL = 0, 1, 2, 3, 4, 5


class BAParams(C.Structure):
    _fields_ = [
        ("huber_delta", C.c_double), ("chi2_thr", C.c_double), ("lambda0", C.c_double),
        ("lambda_up", C.c_double), ("lambda_down", C.c_double), ("lambda_min", C.c_double),
        ("lambda_max", C.c_double), ("accept_rtol", C.c_double), ("max_iter", C.c_int32), ("solver_kind", C.c_int32),
        ("schur_kind", C.c_int32), ("pcg_max_iter", C.c_int32), ("pcg_tol", C.c_double),
        ("use_graph", C.c_int32), ("profile", C.c_int32),
    ]


class BAReport(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_iter", C.c_int32), ("n_accepted", C.c_int32),
        ("solver_used", C.c_int32),
        ("cost_hist", C.c_double * (BA_MAX_ITER + 1)),
        ("cand_cost_hist", C.c_double * BA_MAX_ITER),
        ("lambda_hist", C.c_double * BA_MAX_ITER),
        ("accepted", C.c_uint8 * BA_MAX_ITER),
        ("solve_fail", C.c_int32 * BA_MAX_ITER),
        ("pcg_iters", C.c_int32 * BA_MAX_ITER),
        ("gpu_ms", C.c_float), ("n_launches", C.c_int32),
        ("ms_linearize", C.c_float), ("ms_solve", C.c_float), ("ms_backsub", C.c_float),
        ("ms_other", C.c_float), ("band_width", C.c_int32), ("s_blocks", C.c_int32),
        ("ms_comm", C.c_float), ("ms_point_blocks", C.c_float), ("ms_cam_blocks", C.c_float),
        ("peer_exchange", C.c_int32),
    ]


EXPORTS = [
    "ba_version", "ba_error_string", "ba_last_error", "ba_default_params", "ba_create",
    "ba_destroy", "ba_set_params", "ba_solve_local", "ba_solve_local_host",
    "ba_solve_pose_only", "ba_solve_pose_only_host", "ba_eval_cost", "ba_nccl_unique_id",
    "ba_dist_init", "ba_dist_shutdown", "ba_debug_linearize", "ba_pose_ransac_host",
]

_lib: Optional[C.CDLL] = None


class B200BAError(RuntimeError):
    pass


def lib_path() -> str:
    # BA_LIB_PATH: a diagnostic build of the same library (_build.build(out=..., extra=[...])), never a fallback
    return os.environ.get("BA_LIB_PATH") or _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libb200ba.so (building it with nvcc if absent).  Raises if it cannot be had:
    there is no CPU fallback behind this library."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise B200BAError(f"{path} is missing; run __graft_entry__.build()")
        _build.build()
    lib = C.CDLL(path)
    vp, i32, dp = C.c_void_p, C.c_int32, C.POINTER(C.c_double)
    lib.ba_version.restype = C.c_int
    lib.ba_error_string.restype = C.c_char_p
    lib.ba_error_string.argtypes = [C.c_int]
    lib.ba_last_error.restype = C.c_char_p
    lib.ba_last_error.argtypes = [vp]
    lib.ba_default_params.restype = None
    lib.ba_default_params.argtypes = [C.POINTER(BAParams)]
    lib.ba_create.argtypes = [C.POINTER(BAParams), C.c_int, C.POINTER(vp)]
    lib.ba_destroy.argtypes = [vp]
    lib.ba_set_params.argtypes = [vp, C.POINTER(BAParams)]
    solve_args = [vp, i32, i32, i32, dp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ba_solve_local.argtypes = solve_args + [vp, C.POINTER(BAReport)]
    lib.ba_solve_local_host.argtypes = solve_args + [C.POINTER(BAReport)]
    lib.ba_solve_pose_only.argtypes = [vp, i32, i32, dp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ba_solve_pose_only_host.argtypes = [vp, i32, i32, dp, vp, vp, vp, vp, vp, vp, vp, vp,
                                            C.POINTER(C.c_float)]
    lib.ba_eval_cost.argtypes = [vp, i32, i32, i32, dp, vp, vp, vp, vp, vp, vp, vp, dp]
    lib.ba_pose_ransac_host.argtypes = [vp, i32, dp, vp, vp, i32, C.c_double, C.c_uint32, dp, dp, C.POINTER(i32)]
    lib.ba_nccl_unique_id.argtypes = [vp]
    lib.ba_dist_init.argtypes = [vp, vp, i32, i32]
    lib.ba_dist_shutdown.argtypes = [vp]
    lib.ba_debug_linearize.argtypes = [vp, i32, i32, i32, dp, vp, vp, vp, vp, vp, vp, vp, vp,
                                       C.c_double, vp, vp, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("ba_version",):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(lib: C.CDLL, handle, code: int, what: str):
    if code != BA_OK:
        msg = lib.ba_error_string(code).decode()
        detail = lib.ba_last_error(handle if handle else None).decode()
        raise B200BAError(f"{what}: {msg} ({code}) {detail}")
