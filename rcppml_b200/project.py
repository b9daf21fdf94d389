"""Host mirror of predict() / nnls() / evaluate() for sparse input over the GPU entry points
rcppml_gpu_nnls_double / rcppml_gpu_evaluate_double (ABI extensions, include/rcppml_gpu.h).

Argument meaning follows the reference: predict (R/predict_nmf.R:48 -> Rcpp_predict,
src/RcppFunctions_utils.cpp:23-53), nnls (R/solve.R:84 -> c_nnls, :314-366), evaluate
(R/nmf_methods.R:356 -> Rcpp_evaluate_loss, :152; loss = "mse"). All fp64, like the reference.
There is no CPU fallback: a failing GPU call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _csc(A):
    """Accepts a scipy.sparse matrix or an (indptr, indices, data, shape) tuple."""
    if isinstance(A, tuple):
        indptr, indices, data, shape = A
    else:
        A = A.tocsc()
        A.sort_indices()
        indptr, indices, data, shape = A.indptr, A.indices, A.data, A.shape
    return (np.ascontiguousarray(indptr, np.int32), np.ascontiguousarray(indices, np.int32),
            np.ascontiguousarray(data, np.float64), shape)


def nnls(w, A, *, L1=0.0, L2=0.0, upper_bound=0.0, nonneg=True, cd_maxit=100, cd_tol=1e-8, warm_start=None):
    """Solve A ~ w h for h (k x n), column by column. w: (m, k) like R's `w`. Returns h as (k, n)."""
    lib = _lib.load()
    indptr, indices, data, (m, n) = _csc(A)
    w_T = np.ascontiguousarray(np.asarray(w, np.float64))           # (m, k) C-order == k x m column-major
    if w_T.ndim != 2 or w_T.shape[0] != m:
        raise ValueError(f"nnls: w must have {m} rows (one per row of A)")
    k = w_T.shape[1]
    if warm_start is not None:
        h = np.ascontiguousarray(np.asarray(warm_start, np.float64).T).copy()     # (n, k)
        if h.shape != (n, k):
            raise ValueError(f"nnls: warm_start must be ({k}, {n})")
    else:
        h = np.zeros((n, k), np.float64)
    if indices.size == 0:
        indices, data = np.zeros(1, np.int32), np.zeros(1, np.float64)
    I, D = C.c_int, C.c_double
    st = I(0)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    fn = lib.rcppml_gpu_nnls_double
    fn.restype = None
    fn(ip(indptr), ip(indices), dp(data), C.byref(I(m)), C.byref(I(n)), C.byref(I(int(indptr[n]))), C.byref(I(k)),
       dp(w_T), dp(h), C.byref(D(L1)), C.byref(D(L2)), C.byref(D(upper_bound)), C.byref(I(int(nonneg))),
       C.byref(I(cd_maxit)), C.byref(D(cd_tol)), C.byref(I(int(warm_start is not None))), C.byref(st))
    if st.value != 0:
        raise _lib.NativeLibraryError("rcppml_gpu_nnls_double failed (status != 0); see stderr")
    return h.T.copy()


def predict(w, A, *, L1=0.0, L2=0.0, upper_bound=0.0):
    """Project new samples onto a fixed w: Rcpp_predict(A, mask, w, L1, L2, threads, mask_zeros, upper_bound).
    (The reference accepts `mask`/`mask_zeros` here but never reads them — src/RcppFunctions_utils.cpp:23-53.)"""
    return nnls(w, A, L1=L1, L2=L2, upper_bound=upper_bound, nonneg=True, cd_maxit=100, cd_tol=1e-8)


def evaluate(A, w, d, h, *, mask_zeros=False):
    """Mean squared error of A ~ w diag(d) h. w: (m, k), h: (k, n)."""
    lib = _lib.load()
    indptr, indices, data, (m, n) = _csc(A)
    w_T = np.ascontiguousarray(np.asarray(w, np.float64))
    hh = np.ascontiguousarray(np.asarray(h, np.float64).T)
    dd = np.ascontiguousarray(np.asarray(d, np.float64))
    k = w_T.shape[1]
    if w_T.shape != (m, k) or hh.shape != (n, k) or dd.shape != (k,):
        raise ValueError(f"evaluate: expected w ({m}, k), d (k,), h (k, {n})")
    if indices.size == 0:
        indices, data = np.zeros(1, np.int32), np.zeros(1, np.float64)
    I, D = C.c_int, C.c_double
    st, out = I(0), D(0.0)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    fn = lib.rcppml_gpu_evaluate_double
    fn.restype = None
    fn(ip(indptr), ip(indices), dp(data), C.byref(I(m)), C.byref(I(n)), C.byref(I(int(indptr[n]))), C.byref(I(k)),
       dp(w_T), dp(dd), dp(hh), C.byref(I(int(mask_zeros))), C.byref(out), C.byref(st))
    if st.value != 0:
        raise _lib.NativeLibraryError("rcppml_gpu_evaluate_double failed (status != 0); see stderr")
    return out.value
