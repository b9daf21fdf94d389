"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl
reference`` leg import this module. The product package (rcppml_b200) never does.
See oracle/nmf_oracle.cpp for the restatement and its reference citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RCPPML_ORACLE_LIB: an alternative build of nmf_oracle.cpp (bench.py times an -O3 -march=native build beside the
# package-flags build; the checker itself always uses the default library).
_LIB_PATH = os.environ.get("RCPPML_ORACLE_LIB") or os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nmf_oracle.cpp")
    if os.environ.get("RCPPML_ORACLE_LIB"):
        return _LIB_PATH
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Cfg(C.Structure):
    _fields_ = [
        ("k", C.c_int), ("max_iter", C.c_int), ("tol", C.c_float),
        ("L1_W", C.c_float), ("L1_H", C.c_float), ("L2_W", C.c_float), ("L2_H", C.c_float),
        ("ub_W", C.c_float), ("ub_H", C.c_float),
        ("nonneg_W", C.c_int), ("nonneg_H", C.c_int),
        ("cd_maxit", C.c_int), ("cd_tol", C.c_float),
        ("norm_type", C.c_int), ("solver_mode", C.c_int), ("patience", C.c_int),
        ("threads", C.c_int), ("sort_model", C.c_int), ("has_mask", C.c_int), ("time_budget_s", C.c_double),
    ]


class _Res(C.Structure):
    _fields_ = [
        ("iterations", C.c_int), ("converged", C.c_int), ("train_loss", C.c_float), ("final_tol", C.c_float),
        ("chol_info", C.c_int), ("loop_seconds", C.c_double), ("cd_sweeps", C.c_long),
        ("iter_seconds", C.POINTER(C.c_double)),
    ]


class _CvCfg(C.Structure):
    _fields_ = [
        ("k", C.c_int), ("max_iter", C.c_int), ("tol", C.c_float),
        ("L1_W", C.c_float), ("L1_H", C.c_float), ("L2_W", C.c_float), ("L2_H", C.c_float),
        ("ub_W", C.c_float), ("ub_H", C.c_float), ("nonneg_W", C.c_int), ("nonneg_H", C.c_int),
        ("cd_maxit", C.c_int), ("norm_type", C.c_int), ("solver_mode", C.c_int), ("cv_patience", C.c_int),
        ("threads", C.c_int), ("holdout_fraction", C.c_float), ("cv_seed", C.c_uint32), ("seed", C.c_uint32),
        ("mask_zeros", C.c_int),
    ]


class _CvRes(C.Structure):
    _fields_ = [
        ("iterations", C.c_int), ("converged", C.c_int), ("best_iter", C.c_int),
        ("train_loss", C.c_float), ("test_loss", C.c_float), ("best_test_loss", C.c_float), ("final_tol", C.c_float),
        ("loop_seconds", C.c_double), ("n_test", C.c_long),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_splitmix_hash.restype = C.c_uint64
        _lib.orc_splitmix_hash.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        _lib.orc_is_holdout.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
        _lib.orc_trace_AtA_f32.restype = C.c_float
        _lib.orc_loss_cross_term_f32.restype = C.c_float
        _lib.orc_synth_csc.restype = C.c_long
        _lib.orc_evaluate_mse_f64.restype = C.c_double
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------- RNG ------
def splitmix_next(seed: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.uint64)
    lib().orc_splitmix_next(C.c_uint64(seed), C.c_int(count), _p(out, C.c_uint64))
    return out


def splitmix_hash(seed: int, i: int, j: int) -> int:
    return int(lib().orc_splitmix_hash(seed, i, j))


def is_holdout(seed: int, i: int, j: int, inv_prob: int) -> bool:
    return bool(lib().orc_is_holdout(seed, i, j, inv_prob))


class UniformStream:
    """SplitMix64 sequential stream (rng/rng.hpp:60-104,195-201)."""

    def __init__(self, seed: int):
        self.state = C.c_uint64(12345 if seed == 0 else seed)

    def fill_f32(self, rows: int, cols: int) -> np.ndarray:
        """Column-major rows×cols fill; returned as array of shape (cols, rows) (row = one column)."""
        out = np.empty((cols, rows), dtype=np.float32)
        lib().orc_fill_uniform_f32(C.byref(self.state), _p(out, C.c_float), C.c_long(rows * cols))
        return out

    def fill_f64(self, rows: int, cols: int) -> np.ndarray:
        out = np.empty((cols, rows), dtype=np.float64)
        lib().orc_fill_uniform_f64(C.byref(self.state), _p(out, C.c_double), C.c_long(rows * cols))
        return out


def initialize_factors(k: int, m: int, n: int, seed: int):
    """nmf/nmf_init.hpp:167-182 — W_T then H from ONE stream. Returns (W_T[m,k], H[n,k]) row = factor vector."""
    s = UniformStream(seed)
    return s.fill_f32(k, m), s.fill_f32(k, n)


def init_h_given_w(k: int, n: int, seed: int):
    """nmf/fit_cpu.hpp:203-206 — H from a fresh SplitMix64(seed==0?12345:seed) when W_init is supplied."""
    return UniformStream(seed).fill_f32(k, n)


# ----------------------------------------------------------- primitives ----
# Dense factor matrices are passed as C-contiguous arrays of shape (cols, k):
# row c holds column c of the reference's k×cols column-major matrix.
def gram(F: np.ndarray, threads: int = 1) -> np.ndarray:
    if F.dtype == np.float64:
        F = np.ascontiguousarray(F)
        n, k = F.shape
        G = np.empty((k, k), dtype=np.float64)
        lib().orc_gram_f64(_p(F, C.c_double), k, C.c_long(n), _p(G, C.c_double), threads)
        return G
    F = _f32(F)
    n, k = F.shape
    G = np.empty((k, k), dtype=np.float32)
    lib().orc_gram_f32(_p(F, C.c_float), k, C.c_long(n), _p(G, C.c_float), threads)
    return G


def cd_nnls_col(G, b, x, L1=0.0, L2=0.0, nonneg=True, maxit=100, ub=0.0, cd_tol=0.0):
    """In-place on b and x (same dtype as G). Returns sweeps."""
    k = G.shape[0]
    if G.dtype == np.float64:
        return lib().orc_cd_nnls_col_f64(_p(G, C.c_double), _p(b, C.c_double), _p(x, C.c_double), k,
                                         C.c_double(L1), C.c_double(L2), int(nonneg), maxit, C.c_double(ub),
                                         C.c_double(cd_tol))
    return lib().orc_cd_nnls_col_f32(_p(G, C.c_float), _p(b, C.c_float), _p(x, C.c_float), k, C.c_float(L1),
                                     C.c_float(L2), int(nonneg), maxit, C.c_float(ub), C.c_float(cd_tol))


def nnls_batch_f64(G, B, X, cd_maxit=100, cd_tol=1e-8, L1=0.0, L2=0.0, nonneg=True, ub=0.0, warm_start=False):
    """nnls_batch<CPU,double> (nnls_batch.hpp:150-185). B, X: (n, k) float64, modified in place."""
    n, k = B.shape
    lib().orc_nnls_batch_f64(_p(G, C.c_double), _p(B, C.c_double), _p(X, C.c_double), k, C.c_long(n), cd_maxit,
                             C.c_double(cd_tol), C.c_double(L1), C.c_double(L2), int(nonneg), C.c_double(ub),
                             int(warm_start))


def cholesky_factor(G):
    G = _f32(G)
    k = G.shape[0]
    L = np.empty((k, k), dtype=np.float32)   # column-major k×k stored as (col, row)
    info = lib().orc_cholesky_factor_f32(_p(G, C.c_float), k, _p(L, C.c_float))
    return L, info


def cholesky_solve(L, b):
    x = _f32(b).copy()
    lib().orc_cholesky_solve_f32(_p(L, C.c_float), L.shape[0], _p(x, C.c_float))
    return x


def transpose_csc(Ap, Ai, Ax, m, n):
    Ap, Ai, Ax = _i32(Ap), _i32(Ai), _f32(Ax)
    nnz = int(Ap[n])
    Tp = np.empty(m + 1, dtype=np.int32)
    Ti = np.empty(nnz, dtype=np.int32)
    Tx = np.empty(nnz, dtype=np.float32)
    lib().orc_transpose_csc_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(m), C.c_long(n),
                                _p(Tp, C.c_int), _p(Ti, C.c_int), _p(Tx, C.c_float))
    return Tp, Ti, Tx


def extract_scaling(X, norm_type=0, threads=1):
    """In place on X (cols, k) float32. Returns d."""
    n, k = X.shape
    d = np.empty(k, dtype=np.float32)
    lib().orc_extract_scaling_f32(_p(X, C.c_float), k, C.c_long(n), _p(d, C.c_float), norm_type, threads)
    return d


def half_step(Ap, Ai, Ax, F, G, X, *, solver_mode=0, cd_maxit=100, cd_tol=1e-8, L1=0.0, nonneg=True,
              warm_start=True, ub_in_solver=0.0, threads=1):
    """fused_rhs_nnls_sparse / fused_rhs_cholesky_sparse over all columns; X (n_cols,k) in place."""
    n_cols, k = X.shape
    sw = C.c_long(0)
    info = lib().orc_half_step_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(n_cols),
                                   _p(F, C.c_float), k, _p(G, C.c_float), _p(X, C.c_float), solver_mode, cd_maxit,
                                   C.c_float(cd_tol), C.c_float(L1), int(nonneg), int(warm_start),
                                   C.c_float(ub_in_solver), threads, C.byref(sw))
    return info, sw.value


def rhs(Ap, Ai, Ax, n_cols, F, threads=1):
    k = F.shape[1]
    B = np.empty((n_cols, k), dtype=np.float32)
    lib().orc_rhs_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(n_cols), _p(F, C.c_float), k,
                      _p(B, C.c_float), threads)
    return B


def solve_given_rhs(B, G, X, *, solver_mode=0, cd_maxit=100, cd_tol=1e-8, L1=0.0, nonneg=True, warm_start=True,
                    threads=1):
    n_cols, k = X.shape
    return lib().orc_solve_given_rhs_f32(_p(B, C.c_float), C.c_long(n_cols), k, _p(G, C.c_float), _p(X, C.c_float),
                                         solver_mode, cd_maxit, C.c_float(cd_tol), C.c_float(L1), int(nonneg),
                                         int(warm_start), threads)


def masked_nnls(Ap, Ai, Ax, m_rows, F, G_full, X, Mp, Mi, *, L1=0.0, L2=0.0, nonneg=True, cd_maxit=100,
                cd_tol=1e-8, solver_mode=0, threads=1, warm_start=True):
    n, k = X.shape
    lib().orc_masked_nnls_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(n), C.c_long(m_rows),
                              _p(F, C.c_float), k, _p(G_full, C.c_float), _p(X, C.c_float), _p(Mp, C.c_int),
                              _p(Mi, C.c_int), C.c_float(L1), C.c_float(L2), int(nonneg), cd_maxit,
                              C.c_float(cd_tol), solver_mode, threads, int(warm_start))


def trace_AtA(Ax):
    Ax = _f32(Ax)
    return float(lib().orc_trace_AtA_f32(_p(Ax, C.c_float), C.c_long(Ax.size)))


def loss_cross_term(Atp, Ati, Atx, W_T, H, d, threads=1):
    m, k = W_T.shape
    return float(lib().orc_loss_cross_term_f32(_p(Atp, C.c_int), _p(Ati, C.c_int), _p(Atx, C.c_float),
                                                C.c_long(m), _p(W_T, C.c_float), _p(H, C.c_float),
                                                _p(d, C.c_float), k, threads))


def synth_csc(m, n_local, col_begin=0, density=1e-3, seed=20260101, m_keep=None, threads=0):
    """SURVEY.md §8d generator on the host (OpenMP). Returns (indptr, indices, data)."""
    m_keep = m if m_keep is None else m_keep
    threads = threads or max_threads()
    Ap = np.zeros(n_local + 1, dtype=np.int32)
    nnz = lib().orc_synth_csc(C.c_long(m), C.c_long(n_local), C.c_long(col_begin), C.c_double(density),
                              C.c_uint64(seed), C.c_long(m_keep), 0, _p(Ap, C.c_int), None, None, threads)
    Ai = np.empty(nnz, dtype=np.int32)
    Ax = np.empty(nnz, dtype=np.float32)
    lib().orc_synth_csc(C.c_long(m), C.c_long(n_local), C.c_long(col_begin), C.c_double(density), C.c_uint64(seed),
                        C.c_long(m_keep), 1, _p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), threads)
    return Ap, Ai, Ax


def project_f64(Ap, Ai, Ax, m, n, w_T, *, L1=0.0, L2=0.0, upper_bound=0.0, nonneg=True, cd_maxit=100, cd_tol=1e-8,
                warm_start=None, threads=1):
    """Rcpp_predict / c_nnls (src/RcppFunctions_utils.cpp:23-53, 314-366), fp64. w_T: (m, k). Returns h (n, k)."""
    Ap, Ai = _i32(Ap), _i32(Ai)
    Ax = np.ascontiguousarray(Ax, np.float64)
    w_T = np.ascontiguousarray(w_T, np.float64)
    k = w_T.shape[1]
    h = np.zeros((n, k), np.float64) if warm_start is None else np.array(warm_start, np.float64, order="C")
    lib().orc_project_f64(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_double), C.c_long(m), C.c_long(n), k,
                          _p(w_T, C.c_double), _p(h, C.c_double), C.c_double(L1), C.c_double(L2),
                          C.c_double(upper_bound), int(nonneg), cd_maxit, C.c_double(cd_tol),
                          int(warm_start is not None), threads)
    return h


def evaluate_mse_f64(Ap, Ai, Ax, m, n, w_T, d, h, mask_zeros=False):
    """compute_mse (src/RcppFunctions_utils.cpp:60-90), fp64. w_T: (m, k), h: (n, k)."""
    Ap, Ai = _i32(Ap), _i32(Ai)
    Ax = np.ascontiguousarray(Ax, np.float64)
    w_T = np.ascontiguousarray(w_T, np.float64)
    h = np.ascontiguousarray(h, np.float64)
    d = np.ascontiguousarray(d, np.float64)
    return float(lib().orc_evaluate_mse_f64(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_double), C.c_long(m),
                                            C.c_long(n), w_T.shape[1], _p(w_T, C.c_double), _p(d, C.c_double),
                                            _p(h, C.c_double), int(mask_zeros)))


@dataclass
class OracleCvResult:
    W_T: np.ndarray          # (m, k) normalised
    H: np.ndarray            # (n, k) with d absorbed (fit_cv.hpp:1639-1641)
    d: np.ndarray
    iterations: int
    converged: bool
    best_iter: int
    train_loss: float
    test_loss: float
    best_test_loss: float
    final_tol: float
    n_test: int
    train_history: np.ndarray
    test_history: np.ndarray
    loop_seconds: float


def nmf_fit_cv(Ap, Ai, Ax, m, n, k, W_T0, H0, *, max_iter=100, tol=1e-4, L1=(0.0, 0.0), L2=(0.0, 0.0),
               upper_bound=(0.0, 0.0), nonneg=(True, True), cd_maxit=100, norm_type=0, solver_mode=0, cv_patience=5,
               threads=0, holdout_fraction=0.1, cv_seed=0, seed=42, mask_zeros=True) -> OracleCvResult:
    """nmf_fit_cv<CPU,float,Sparse> (nmf/fit_cv.hpp:124), MSE / standard variant / no user mask."""
    Ap, Ai, Ax = _i32(Ap), _i32(Ai), _f32(Ax)
    W_T, H = _f32(W_T0).copy(), _f32(H0).copy()
    d = np.ones(k, np.float32)
    cfg = _CvCfg(k=k, max_iter=max_iter, tol=tol, L1_W=L1[0], L1_H=L1[1], L2_W=L2[0], L2_H=L2[1],
                 ub_W=upper_bound[0], ub_H=upper_bound[1], nonneg_W=int(nonneg[0]), nonneg_H=int(nonneg[1]),
                 cd_maxit=cd_maxit, norm_type=norm_type, solver_mode=solver_mode, cv_patience=cv_patience,
                 threads=threads, holdout_fraction=holdout_fraction, cv_seed=cv_seed, seed=seed,
                 mask_zeros=int(mask_zeros))
    tr, te = np.zeros(max_iter, np.float32), np.zeros(max_iter, np.float32)
    res = _CvRes()
    rc = lib().orc_nmf_fit_cv_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(m), C.c_long(n),
                                  C.byref(cfg), _p(W_T, C.c_float), _p(H, C.c_float), _p(d, C.c_float),
                                  _p(tr, C.c_float), _p(te, C.c_float), C.byref(res))
    if rc != 0:
        raise ValueError("oracle: invalid CV configuration")
    it = res.iterations
    return OracleCvResult(W_T, H, d, it, bool(res.converged), res.best_iter, res.train_loss, res.test_loss,
                          res.best_test_loss, res.final_tol, res.n_test, tr[:it].copy(), te[:it].copy(),
                          res.loop_seconds)


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ------------------------------------------------------------- full fit ----
@dataclass
class OracleResult:
    W_T: np.ndarray          # (m, k): row ℓ = factor vector of row ℓ of A (normalised)
    H: np.ndarray            # (n, k): row j = column j of the reference's k×n H
    d: np.ndarray
    iterations: int
    converged: bool
    train_loss: float
    final_tol: float
    chol_info: int
    loop_seconds: float
    cd_sweeps: int
    loss_history: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    iter_seconds: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float64))


def nmf_fit(Ap, Ai, Ax, m, n, k, W_T0, H0, *, max_iter=100, tol=1e-4, L1=(0.0, 0.0), L2=(0.0, 0.0),
            upper_bound=(0.0, 0.0), nonneg=(True, True), cd_maxit=100, cd_tol=1e-8, norm_type=0, solver_mode=0,
            patience=5, threads=0, sort_model=False, mask=None, time_budget_s=0.0) -> OracleResult:
    """nmf_fit<CPU,float,Sparse> (nmf/fit_cpu.hpp:172). Pairs are (W, H) as at the R boundary
    (src/RcppFunctions_nmf.cpp:59-72). mask = (Mp, Mi) CSC pattern of masked entries or None."""
    Ap, Ai, Ax = _i32(Ap), _i32(Ai), _f32(Ax)
    W_T = _f32(W_T0).copy()
    H = _f32(H0).copy()
    assert W_T.shape == (m, k) and H.shape == (n, k)
    d = np.ones(k, dtype=np.float32)
    cfg = _Cfg(k=k, max_iter=max_iter, tol=tol, L1_W=L1[0], L1_H=L1[1], L2_W=L2[0], L2_H=L2[1],
               ub_W=upper_bound[0], ub_H=upper_bound[1], nonneg_W=int(nonneg[0]), nonneg_H=int(nonneg[1]),
               cd_maxit=cd_maxit, cd_tol=cd_tol, norm_type=norm_type, solver_mode=solver_mode, patience=patience,
               threads=threads, sort_model=int(sort_model), has_mask=int(mask is not None),
               time_budget_s=time_budget_s)
    Mp = Mi = None
    if mask is not None:
        Mp, Mi = _i32(mask[0]), _i32(mask[1])
    hist = np.zeros(max_iter, dtype=np.float32)
    res = _Res()
    it_s = np.zeros(max_iter, dtype=np.float64)
    res.iter_seconds = _p(it_s, C.c_double)
    rc = lib().orc_nmf_fit_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(m), C.c_long(n),
                               C.byref(cfg), _p(W_T, C.c_float), _p(H, C.c_float), _p(d, C.c_float),
                               _p(Mp, C.c_int), _p(Mi, C.c_int), _p(hist, C.c_float), C.byref(res))
    if rc != 0:
        raise ValueError("oracle: invalid configuration (core/config.hpp:421-432)")
    return OracleResult(W_T, H, d, res.iterations, bool(res.converged), res.train_loss, res.final_tol,
                        res.chol_info, res.loop_seconds, res.cd_sweeps, hist[:res.iterations].copy(),
                        it_s[:res.iterations].copy())
