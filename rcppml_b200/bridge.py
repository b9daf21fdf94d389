"""ctypes twin of the reference's CPU-side dlsym bridge (inst/include/FactorNet/gpu/bridge_nmf.hpp).

`bridge_nmf_sparse` packs its arguments exactly as bridge_nmf.hpp:199-342 does — CSC indices as
int32, values / W / H / d as double, every scalar behind a pointer, non-null length-1 dummies
for the unused graph / guide arrays — and calls `rcppml_gpu_nmf_unified_float` out of
RcppML_gpu.so, i.e. it exercises the drop-in boundary the way the R package would.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class BridgeResult:
    W_T: np.ndarray       # (m, k) — row ℓ = column ℓ of the k×m wire matrix
    H: np.ndarray         # (n, k)
    d: np.ndarray
    iterations: int
    converged: bool
    train_loss: float
    final_tol: float
    status: int


def gpu_detect(max_gpus: int = 8):
    """R/gpu_backend.R:101-106 `.C("rcppml_gpu_detect", ...)`."""
    lib = _lib.load()
    n = C.c_int(0)
    st = C.c_int(0)
    mx = C.c_int(max_gpus)
    total = (C.c_double * max_gpus)()
    free = (C.c_double * max_gpus)()
    lib.rcppml_gpu_detect(C.byref(n), total, free, C.byref(mx), C.byref(st))
    cnt = min(n.value, max_gpus)
    return {"num_gpus": n.value, "status": st.value, "total_mem_mb": list(total)[:cnt],
            "free_mem_mb": list(free)[:cnt]}


class PackedCall:
    """The 73 pointer arguments of rcppml_gpu_nmf_unified_float, packed once (bridge_nmf.hpp:199-307).

    col_ptr/row_idx must be int32, values/W/H float64 C-contiguous; they are passed BY REFERENCE
    (W, H are overwritten in place, like the reference's W_flat/H_flat), so a caller can keep them in
    pinned memory and time exactly the native call.
    """

    def __init__(self, col_ptr, row_idx, values, m, n, k, W, H, *, max_iter=100, tol=1e-4, L1=(0.0, 0.0),
                 L2=(0.0, 0.0), L21=(0.0, 0.0), angular=(0.0, 0.0), upper_bound=(0.0, 0.0), nonneg=(True, True),
                 cd_maxit=100, verbose=False, seed=42, loss_every=1, patience=5, loss_type=0, norm_type=0,
                 projective=False, symmetric=False, solver_mode=0, mask=None):
        lib = _lib.load()
        if col_ptr.dtype != np.int32 or row_idx.dtype != np.int32:
            raise TypeError("PackedCall: col_ptr / row_idx must be int32 (the wire format, bridge_nmf.hpp:39-41)")
        if values.dtype != np.float64 or W.dtype != np.float64 or H.dtype != np.float64:
            raise TypeError("PackedCall: values / W / H must be float64 (the wire format)")
        if W.shape != (m, k) or H.shape != (n, k) or not (W.flags.c_contiguous and H.flags.c_contiguous):
            raise ValueError(f"PackedCall: W must be C-contiguous ({m}, {k}) and H ({n}, {k})")
        nnz = int(col_ptr[n])
        self.W, self.H = W, H
        self.d = np.ones(k, dtype=np.float64)                                 # :208
        I, D = C.c_int, C.c_double
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self._keep = [col_ptr, row_idx, values]

        def i_(v):
            x = I(int(v)); self._keep.append(x); return C.byref(x)

        def d_(v):
            x = D(float(v)); self._keep.append(x); return C.byref(x)

        one_i = np.zeros(1, np.int32)          # dummies: pack_graph :144-151, guides :300-304
        one_d = np.zeros(1, np.float64)
        theta = np.zeros(max(m, 1), np.float64)                               # :284
        self._keep += [one_i, one_d, theta]
        self._theta_len, self._iter, self._conv, self._status = I(0), I(0), I(0), I(0)
        self._loss, self._tol = D(0.0), D(0.0)
        self._args = [
            ip(col_ptr), ip(row_idx), dp(values),
            i_(m), i_(n), i_(nnz), i_(k),
            dp(W), dp(H), dp(self.d),
            i_(max_iter), d_(tol),
            d_(L1[1]), d_(L1[0]), d_(L2[1]), d_(L2[0]),          # L1_H, L1_W, L2_H, L2_W
            d_(L21[1]), d_(L21[0]),
            d_(angular[1]), d_(angular[0]),
            d_(upper_bound[1]), d_(upper_bound[0]),
            i_(cd_maxit), i_(verbose), i_(seed),
            i_(loss_every), i_(patience),
            i_(nonneg[0]), i_(nonneg[1]),
            i_(loss_type), d_(1.0),
            i_(20), d_(1e-4),
            i_(norm_type),
            i_(projective), i_(symmetric),
            i_(solver_mode),
            ip(one_i), ip(one_i), dp(one_d), i_(0), i_(0), d_(0.0),       # graph_W
            ip(one_i), ip(one_i), dp(one_d), i_(0), i_(0), d_(0.0),       # graph_H
            i_(0),
            d_(0.1), d_(5.0), d_(0.0), d_(10.0), d_(1e6), d_(1e-2), d_(1.0), d_(1e4), d_(1e-6),
            d_(0.0), d_(1.5),
            dp(theta), C.byref(self._theta_len),
            ip(one_i), ip(one_i), dp(one_d), ip(one_i), i_(0),
            C.byref(self._iter), C.byref(self._conv), C.byref(self._loss),
            C.byref(self._status),
            C.byref(self._tol),
        ]
        assert len(self._args) == 73, len(self._args)
        self._fn = lib.rcppml_gpu_nmf_unified_float
        if mask is not None:                  # ABI extension: mask pattern + the same 73 arguments
            mp = np.ascontiguousarray(mask[0], np.int32)
            mi = np.ascontiguousarray(mask[1], np.int32)
            if mi.size == 0:
                mi = np.zeros(1, np.int32)
            self._keep += [mp, mi]
            self._args = [ip(mp), ip(mi), i_(int(mp[n]))] + self._args
            self._fn = lib.rcppml_gpu_nmf_masked_unified_float
        self._fn.restype = None

    def __call__(self):
        self._fn(*self._args)
        return self

    status = property(lambda self: self._status.value)
    iterations = property(lambda self: self._iter.value)
    converged = property(lambda self: bool(self._conv.value))
    train_loss = property(lambda self: self._loss.value)
    final_tol = property(lambda self: self._tol.value)


def bridge_nmf_sparse(indptr, indices, data, m, n, k, W_T0, H0, **kw) -> BridgeResult:
    """All (a, b) keyword pairs are (W, H). W_T0: (m, k), H0: (n, k) — any float dtype (sent as double)."""
    col_ptr = np.ascontiguousarray(indptr, dtype=np.int32)
    row_idx = np.ascontiguousarray(indices, dtype=np.int32)
    values = np.ascontiguousarray(data, dtype=np.float64)                 # bridge_nmf.hpp:200-203
    if row_idx.size == 0:
        row_idx = np.zeros(1, np.int32)
        values = np.zeros(1, np.float64)
    W = np.array(W_T0, dtype=np.float64, order="C")                       # :206-219 (k×m col-major)
    H = np.array(H0, dtype=np.float64, order="C")                         # :221-230
    call = PackedCall(col_ptr, row_idx, values, m, n, k, W, H, **kw)()
    return BridgeResult(W.astype(np.float32), H.astype(np.float32), call.d.astype(np.float32), call.iterations,
                        call.converged, call.train_loss, call.final_tol, call.status)


@dataclass
class BridgeCvResult:
    W_T: np.ndarray
    H: np.ndarray            # d absorbed (fit_cv.hpp:1639-1646)
    d: np.ndarray
    iterations: int
    converged: bool
    train_loss: float
    test_loss: float
    best_test_loss: float
    best_iter: int
    status: int


def bridge_nmf_cv_sparse(indptr, indices, data, m, n, k, W_T0, H0, *, max_iter=100, tol=1e-4, L1=(0.0, 0.0),
                         L2=(0.0, 0.0), nonneg=(True, True), cd_maxit=100, verbose=False, seed=42, holdout_fraction=0.1,
                         cv_seed=0, mask_zeros=True, norm_type=0, loss_type=0, projective=False, symmetric=False,
                         solver_mode=0) -> BridgeCvResult:
    """ctypes twin of bridge_nmf_cv_sparse (gpu/bridge_nmf.hpp:399+): the 51-pointer
    rcppml_gpu_nmf_cv_unified_float call (type at gpu/bridge_nmf.hpp:78-99). Pairs are (W, H)."""
    lib = _lib.load()
    col_ptr = np.ascontiguousarray(indptr, dtype=np.int32)
    row_idx = np.ascontiguousarray(indices, dtype=np.int32)
    values = np.ascontiguousarray(data, dtype=np.float64)
    W = np.array(W_T0, dtype=np.float64, order="C")
    H = np.array(H0, dtype=np.float64, order="C")
    d = np.ones(k, dtype=np.float64)
    I, D = C.c_int, C.c_double
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    keep = []

    def i_(v):
        x = I(int(v)); keep.append(x); return C.byref(x)

    def d_(v):
        x = D(float(v)); keep.append(x); return C.byref(x)

    one_i, one_d = np.zeros(1, np.int32), np.zeros(1, np.float64)
    out_iter, out_conv, out_best_iter, out_status = I(0), I(0), I(0), I(0)
    out_train, out_test, out_best = D(0.0), D(0.0), D(0.0)
    args = [
        ip(col_ptr), ip(row_idx), dp(values),
        i_(m), i_(n), i_(int(col_ptr[n])), i_(k),
        dp(W), dp(H), dp(d),
        i_(max_iter), d_(tol),
        d_(L1[1]), d_(L1[0]), d_(L2[1]), d_(L2[0]),
        i_(cd_maxit), i_(verbose), i_(seed),
        d_(holdout_fraction), i_(cv_seed), i_(mask_zeros),
        i_(nonneg[0]), i_(nonneg[1]),
        i_(norm_type),
        i_(loss_type), d_(1.0),
        i_(20), d_(1e-4),
        ip(one_i), ip(one_i), dp(one_d), i_(0), i_(0), d_(0.0),
        ip(one_i), ip(one_i), dp(one_d), i_(0), i_(0), d_(0.0),
        i_(projective), i_(symmetric), i_(solver_mode),
        C.byref(out_iter), C.byref(out_conv),
        C.byref(out_train), C.byref(out_test),
        C.byref(out_best), C.byref(out_best_iter),
        C.byref(out_status),
    ]
    assert len(args) == 51, len(args)
    fn = lib.rcppml_gpu_nmf_cv_unified_float
    fn.restype = None
    fn(*args)
    return BridgeCvResult(W.astype(np.float32), H.astype(np.float32), d.astype(np.float32), out_iter.value,
                          bool(out_conv.value), out_train.value, out_test.value, out_best.value, out_best_iter.value,
                          out_status.value)


def gpu_nmf_zerocopy(col_ptr_addr: int, row_idx_addr: int, values_addr: int, m, n, nnz, k, W_T0, H0, *, maxit=100,
                     tol=1e-4, seed=42, L1=(0.0, 0.0), L2=(0.0, 0.0), L21=(0.0, 0.0), ortho=(0.0, 0.0),
                     upper_bound=(0.0, 0.0), cd_maxit=10, verbose=False, nonneg=(True, True), loss_every=1, patience=5,
                     loss_type=0, huber_delta=1.0, irls_max_iter=20, irls_tol=1e-4, norm_type=0) -> BridgeResult:
    """ctypes twin of `.gpu_nmf_zerocopy` (R/gpu_backend.R:183-265): the 39-argument `.C` call of
    rcppml_gpu_nmf_zerocopy_double. The CSC arrays (int32 col_ptr / row_idx, float64 values) are DEVICE arrays
    whose addresses travel as doubles. NB this R wrapper sends the FIRST element of each penalty pair as the
    H value (L1_H = L1[1] in R's 1-based indexing, R/gpu_backend.R:240-249) while nonneg is (W, H); kept."""
    lib = _lib.load()
    W = np.array(W_T0, dtype=np.float64, order="C")
    H = np.array(H0, dtype=np.float64, order="C")
    if W.shape != (m, k) or H.shape != (n, k):
        raise ValueError(f"gpu_nmf_zerocopy: W_T0 must be ({m}, {k}) and H0 ({n}, {k})")
    d = np.ones(k, dtype=np.float64)
    I, D = C.c_int, C.c_double
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    keep = []

    def i_(v):
        x = I(int(v)); keep.append(x); return C.byref(x)

    def d_(v):
        x = D(float(v)); keep.append(x); return C.byref(x)

    out_iter, out_conv, out_status = I(0), I(0), I(0)
    out_loss, out_tol = D(0.0), D(0.0)
    args = [
        d_(col_ptr_addr), d_(row_idx_addr), d_(values_addr),
        i_(m), i_(n), d_(nnz), i_(k),
        dp(W), dp(H), dp(d),
        i_(maxit), d_(tol),
        d_(L1[0]), d_(L1[1]), d_(L2[0]), d_(L2[1]),              # *_H = pair[1] in R = first element
        d_(L21[0]), d_(L21[1]),
        d_(ortho[0]), d_(ortho[1]),
        d_(upper_bound[0]), d_(upper_bound[1]),
        i_(cd_maxit), i_(verbose), i_(seed),
        i_(loss_every), i_(patience),
        i_(nonneg[0]), i_(nonneg[1]),
        i_(loss_type), d_(huber_delta),
        i_(irls_max_iter), d_(irls_tol),
        i_(norm_type),
        C.byref(out_iter), C.byref(out_conv), C.byref(out_loss),
        C.byref(out_status),
        C.byref(out_tol),
    ]
    assert len(args) == 39, len(args)
    fn = lib.rcppml_gpu_nmf_zerocopy_double
    fn.restype = None
    fn(*args)
    return BridgeResult(W.astype(np.float32), H.astype(np.float32), d.astype(np.float32), out_iter.value,
                        bool(out_conv.value), out_loss.value, out_tol.value, out_status.value)
