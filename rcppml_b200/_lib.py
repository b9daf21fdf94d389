"""Loader of the native library (rcppml_b200/lib/RcppML_gpu.so).

There is no Python/CPU fallback: if the CUDA library is missing or a call fails, the error is
raised to the caller (the reference gateway owns the CPU path, nmf/fit.hpp:125-133).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RCPPML_B200_LIB selects an experiment build (csrc/Makefile VARIANT=...); default is the product library.
LIB_PATH = os.environ.get("RCPPML_B200_LIB") or os.path.join(_HERE, "lib", "RcppML_gpu.so")

NUM_SECTIONS = 8
SECTION_NAMES = ("gram_H", "fused_rhs_nnls_H", "scaling_H", "gram_W", "fused_rhs_nnls_W", "scaling_W", "loss",
                 "comm")


class NativeLibraryError(RuntimeError):
    pass


class Config(C.Structure):
    """rcppml_b200_config (include/rcppml_gpu.h)."""
    _fields_ = [
        ("k", C.c_int), ("max_iter", C.c_int), ("tol", C.c_float),
        ("L1_W", C.c_float), ("L1_H", C.c_float), ("L2_W", C.c_float), ("L2_H", C.c_float),
        ("ub_W", C.c_float), ("ub_H", C.c_float),
        ("nonneg_W", C.c_int), ("nonneg_H", C.c_int),
        ("cd_maxit", C.c_int), ("cd_tol", C.c_float),
        ("norm_type", C.c_int), ("solver_mode", C.c_int), ("patience", C.c_int), ("verbose", C.c_int),
    ]


class Result(C.Structure):
    """rcppml_b200_result (include/rcppml_gpu.h)."""
    _fields_ = [
        ("iterations", C.c_int), ("converged", C.c_int), ("train_loss", C.c_float), ("final_tol", C.c_float),
        ("status", C.c_int), ("gpu_launches", C.c_int), ("loop_ms", C.c_double),
    ]


class CvConfig(C.Structure):
    """rcppml_b200_cv_config (include/rcppml_gpu.h)."""
    _fields_ = [("holdout_fraction", C.c_float), ("cv_seed", C.c_uint32), ("seed", C.c_uint32),
                ("mask_zeros", C.c_int), ("cv_patience", C.c_int)]


class CvResult(C.Structure):
    """rcppml_b200_cv_result (include/rcppml_gpu.h)."""
    _fields_ = [("train_loss", C.c_float), ("test_loss", C.c_float), ("best_test_loss", C.c_float),
                ("best_iter", C.c_int), ("n_test", C.c_int64)]


_lib = None


def _preload_nccl():
    # The library's DT_NEEDED libnccl.so.2 resolves through its rpath (torch-bundled NCCL); when
    # torch is already imported the same soname is already mapped and is reused.
    try:
        import torch  # noqa: F401  (maps libnccl.so.2 and libcudart)
    except Exception:
        pass


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(rcppml_b200 has no CPU fallback)")
    _preload_nccl()
    try:
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    except OSError as ex:
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {ex}") from ex
    E = C.c_void_p
    lib.rcppml_b200_last_error.restype = C.c_char_p
    lib.rcppml_b200_engine_create.argtypes = [C.POINTER(E), C.c_int]
    lib.rcppml_b200_engine_destroy.argtypes = [E]
    lib.rcppml_b200_engine_destroy.restype = None
    ip, fp, dp = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    lib.rcppml_b200_set_matrix_f32.argtypes = [E, C.c_int, C.c_int, C.c_int64, ip, ip, fp]
    lib.rcppml_b200_set_matrix_f64.argtypes = [E, C.c_int, C.c_int, C.c_int64, ip, ip, dp]
    lib.rcppml_b200_set_matrix_with_transpose_f32.argtypes = [E, C.c_int, C.c_int, C.c_int64, ip, ip, fp, ip, ip, fp]
    lib.rcppml_b200_set_matrix_with_transpose_f64.argtypes = [E, C.c_int, C.c_int, C.c_int64, ip, ip, dp, ip, ip, dp]
    lib.rcppml_b200_set_matrix_synthetic.argtypes = [E, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64]
    lib.rcppml_b200_set_matrix_synthetic_sharded.argtypes = [E, C.c_int, C.c_int, C.c_double, C.c_uint64]
    lib.rcppml_b200_set_matrix_sharded_f32.argtypes = [E, C.c_int, C.c_int, ip, ip, fp, ip, ip, fp]
    lib.rcppml_b200_get_shard.argtypes = [E, ip, ip, ip, ip, C.POINTER(C.c_int64)]
    lib.rcppml_b200_set_partition.argtypes = [E, ip, ip]
    lib.rcppml_b200_factor_checksum.argtypes = [E, C.POINTER(C.c_uint64)]
    lib.rcppml_b200_get_matrix.argtypes = [E, C.POINTER(C.c_int64), ip, ip, fp]
    lib.rcppml_b200_get_matrix_t.argtypes = [E, ip, ip, fp]
    lib.rcppml_b200_set_mask.argtypes = [E, C.c_int64, ip, ip]
    lib.rcppml_b200_set_factors_f32.argtypes = [E, C.c_int, fp, fp]
    lib.rcppml_b200_set_factors_f64.argtypes = [E, C.c_int, dp, dp]
    lib.rcppml_b200_init_factors.argtypes = [E, C.c_int, C.c_uint32, C.c_int]
    lib.rcppml_b200_get_factors_f32.argtypes = [E, fp, fp, fp]
    lib.rcppml_b200_get_factors_f64.argtypes = [E, dp, dp, dp]
    lib.rcppml_b200_set_factor_blocks_f32.argtypes = [E, C.c_int, fp, fp]
    lib.rcppml_b200_get_factor_blocks_f32.argtypes = [E, fp, fp, fp]
    lib.rcppml_b200_begin_fit.argtypes = [E, C.POINTER(Config)]
    lib.rcppml_b200_iterate.argtypes = [E, C.c_int]
    lib.rcppml_b200_fit.argtypes = [E, C.POINTER(Config)]
    lib.rcppml_b200_fit_cv.argtypes = [E, C.POINTER(Config), C.POINTER(CvConfig)]
    lib.rcppml_b200_get_cv_result.argtypes = [E, C.POINTER(CvResult)]
    lib.rcppml_b200_get_cv_history.argtypes = [E, fp, fp, C.c_int]
    lib.rcppml_b200_get_result.argtypes = [E, C.POINTER(Result)]
    lib.rcppml_b200_get_loss_history.argtypes = [E, fp, C.c_int]
    lib.rcppml_b200_set_profiling.argtypes = [E, C.c_int]
    lib.rcppml_b200_get_profile.argtypes = [E, dp, ip]
    lib.rcppml_b200_half_step.argtypes = [E, C.POINTER(Config), C.c_int, C.c_int, C.c_int]
    lib.rcppml_b200_cd_sweeps.argtypes = [E]
    lib.rcppml_b200_cd_sweeps.restype = C.c_int64
    lib.rcppml_b200_selftest_division.argtypes = [C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]
    lib.rcppml_b200_get_counters.argtypes = [E, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.rcppml_b200_balanced_col_cuts.argtypes = [ip, C.c_int, C.c_int, C.c_int, ip]
    lib.rcppml_b200_last_call_wall_ms.restype = C.c_double
    lib.rcppml_b200_last_call_wall_ms.argtypes = []
    lib.rcppml_b200_nccl_unique_id.argtypes = [C.c_char_p]
    lib.rcppml_b200_comm_init.argtypes = [E, C.c_int, C.c_int, C.c_char_p]
    lib.rcppml_b200_comm_ipc_export.argtypes = [E, C.c_char_p]
    lib.rcppml_b200_comm_ipc_import.argtypes = [E, C.c_char_p]
    for name in ("rcppml_b200_comm_mc_wanted", "rcppml_b200_comm_mc_ready", "rcppml_b200_comm_mc_finish", "rcppml_b200_comm_p2p_close",
                 "rcppml_b200_comm_mc_bind", "rcppml_b200_comm_mc_disable"):
        getattr(lib, name).argtypes = [E]
    lib.rcppml_b200_comm_mc_export.argtypes = [E, C.c_char_p]
    lib.rcppml_b200_comm_mc_import.argtypes = [E, C.c_char_p]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rcppml_b200_last_error()
        raise NativeLibraryError(f"{what} failed: {msg.decode() if msg else 'unknown error'}")
