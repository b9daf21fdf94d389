"""The drop-in boundary exercised by the REFERENCE'S OWN caller.

oracle/_ref/libref_bridge.so is the reference's gpu/bridge_nmf.hpp (bridge_nmf_sparse / bridge_nmf_cv_sparse — the code
that packs the 73 / 51 pointer arguments) and gpu/loader.hpp (detect_gpus_via_bridge), compiled unmodified from
/root/reference against the Eigen stand-in (`make -C oracle ref_hotpath`). It finds the entry points with
dlsym(RTLD_DEFAULT, ...), so the tests load rcppml_b200/lib/RcppML_gpu.so with RTLD_GLOBAL first — what R's
dyn.load(path, local = FALSE) does (R/gpu_backend.R:85-88). A mistake in the argument order of the C ABI that this
repository's own ctypes twin (rcppml_b200/bridge.py) happened to share would show up here."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import RTOL, random_csc, rel_err, zero_pattern_equal

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_BRIDGE = os.path.join(_ROOT, "oracle", "_ref", "libref_bridge.so")
_PRODUCT = os.path.join(_ROOT, "rcppml_b200", "lib", "RcppML_gpu.so")
pytestmark = pytest.mark.skipif(not (os.path.exists(_BRIDGE) and os.path.exists(_PRODUCT)),
                                reason="oracle/_ref/libref_bridge.so or the product library not built")


class Params(C.Structure):
    _fields_ = [("k", C.c_int), ("max_iter", C.c_int), ("tol", C.c_float),
                ("L1_W", C.c_float), ("L1_H", C.c_float), ("L2_W", C.c_float), ("L2_H", C.c_float),
                ("ub_W", C.c_float), ("ub_H", C.c_float),
                ("nonneg_W", C.c_int), ("nonneg_H", C.c_int), ("cd_maxit", C.c_int), ("norm_type", C.c_int),
                ("solver_mode", C.c_int), ("seed", C.c_uint),
                ("holdout_fraction", C.c_float), ("cv_seed", C.c_uint), ("mask_zeros", C.c_int)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("train_loss", C.c_float), ("final_tol", C.c_float),
                ("test_loss", C.c_float), ("best_test_loss", C.c_float), ("best_iter", C.c_int)]


@pytest.fixture(scope="module")
def refbridge():
    C.CDLL(_PRODUCT, mode=C.RTLD_GLOBAL)            # dyn.load(path, local = FALSE)
    return C.CDLL(_BRIDGE)


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def _call(lib, A, k, W0, H0, **kw):
    """W0: (m, k) rows = factor vectors; H0: (n, k). Returns (rc, W (m,k), H (n,k), d, Result, message)."""
    m, n = A.shape
    q = Params(k=k, max_iter=kw.get("max_iter", 5), tol=kw.get("tol", 0.0), L1_W=kw.get("L1", (0, 0))[0],
               L1_H=kw.get("L1", (0, 0))[1], L2_W=kw.get("L2", (0, 0))[0], L2_H=kw.get("L2", (0, 0))[1],
               ub_W=0.0, ub_H=0.0, nonneg_W=1, nonneg_H=1, cd_maxit=kw.get("cd_maxit", 100), norm_type=0,
               solver_mode=kw.get("solver_mode", 0), seed=42, holdout_fraction=0.0, cv_seed=0, mask_zeros=1)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    W_in = np.ascontiguousarray(W0.T)                # m x k column-major = (k, m) C-order
    H_in = np.ascontiguousarray(H0)                  # k x n column-major = (n, k) C-order
    W_out, H_out, d = np.zeros((k, m), np.float32), np.zeros((n, k), np.float32), np.zeros(k, np.float32)
    res, err = Result(), C.create_string_buffer(256)
    rc = lib.refbridge_nmf_sparse_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), m, n, C.byref(q),
                                      _p(W_in, C.c_float), _p(H_in, C.c_float), _p(W_out, C.c_float),
                                      _p(H_out, C.c_float), _p(d, C.c_float), C.byref(res), err, 256)
    return rc, W_out.T.copy(), H_out, d, res, err.value.decode()


def test_reference_bridge_reaches_the_library_and_fails_cleanly_without_a_gpu(refbridge):
    """CPU: the reference's resolve<>() finds our symbols; without a device the library answers status -1 (never a
    CPU fallback), the reference bridge turns that into its runtime_error, and detection reports no GPU plan."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test below")
    cnt, mem = C.c_int(-1), C.c_double(-1.0)
    assert refbridge.refbridge_detect(C.byref(cnt), C.byref(mem)) == 0
    A = random_csc(60, 40, 0.2, 1)
    rng = np.random.default_rng(0)
    rc, *_rest, msg = _call(refbridge, A, 4, rng.random((60, 4)).astype(np.float32), rng.random((40, 4)).astype(np.float32))
    assert rc == -1 and "status != 0" in msg


@pytest.mark.gpu
@pytest.mark.parametrize("k,solver,kw", [(8, 0, {}), (20, 0, dict(L1=(0.01, 0.02))), (64, 1, dict(L2=(0.01, 0.0))),
                                         (32, 1, dict(L1=(0.01, 0.01), L2=(0.01, 0.01)))])
def test_reference_bridge_drives_our_library(refbridge, oracle, k, solver, kw):
    """GPU: gpu/bridge_nmf.hpp::bridge_nmf_sparse -> dlsym -> rcppml_gpu_nmf_unified_float. Same factors as this
    repository's own twin of the bridge (bit for bit) and as the oracle (1e-5)."""
    import rcppml_b200 as rb
    cnt, mem = C.c_int(0), C.c_double(0.0)
    assert refbridge.refbridge_detect(C.byref(cnt), C.byref(mem)) == 1 and cnt.value >= 1 and mem.value > 1e11
    m, n, iters = 700, 450, 6
    A = random_csc(m, n, 0.05, 60 + k, ragged=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    rc, W, H, d, res, msg = _call(refbridge, A, k, W0, H0, max_iter=iters, solver_mode=solver, **kw)
    assert rc == 0, msg
    assert res.iterations == iters
    twin = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver,
                                **kw)
    assert np.array_equal(W, twin.W_T.astype(np.float32)) and np.array_equal(H, twin.H.astype(np.float32))
    assert np.array_equal(d, twin.d.astype(np.float32))
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver, **kw)
    assert rel_err(W, ref.W_T) <= RTOL and rel_err(H, ref.H) <= RTOL and rel_err(d, ref.d) <= RTOL
    assert zero_pattern_equal(W, ref.W_T) and zero_pattern_equal(H, ref.H)
    assert abs(res.train_loss - ref.train_loss) <= 1e-5 * abs(ref.train_loss)


@pytest.mark.gpu
def test_reference_cv_bridge_drives_our_library(refbridge, oracle):
    """GPU: bridge_nmf_cv_sparse -> rcppml_gpu_nmf_cv_unified_float (51 pointers). The bridge draws H itself from
    SplitMix64(seed + 0x9E3779B9) as doubles (bridge_nmf.hpp:226-229); the oracle gets the same H."""
    m, n, k, iters = 500, 300, 8, 5
    A = random_csc(m, n, 0.08, 5, ragged=True)
    W0, _ = oracle.initialize_factors(k, m, n, 42)
    H0 = oracle.UniformStream((42 + 0x9E3779B9) & 0xFFFFFFFFFFFFFFFF).fill_f64(k, n).reshape(n, k).astype(np.float32)
    q = Params(k=k, max_iter=iters, tol=0.0, L1_W=0.0, L1_H=0.0, L2_W=0.0, L2_H=0.0, ub_W=0.0, ub_H=0.0, nonneg_W=1,
               nonneg_H=1, cd_maxit=20, norm_type=0, solver_mode=1, seed=42, holdout_fraction=0.1, cv_seed=7, mask_zeros=1)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    W_in = np.ascontiguousarray(W0.T)
    W_out, H_out, d = np.zeros((k, m), np.float32), np.zeros((n, k), np.float32), np.zeros(k, np.float32)
    res, err = Result(), C.create_string_buffer(256)
    rc = refbridge.refbridge_nmf_cv_sparse_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), m, n, C.byref(q),
                                               _p(W_in, C.c_float), _p(W_out, C.c_float), _p(H_out, C.c_float),
                                               _p(d, C.c_float), C.byref(res), err, 256)
    assert rc == 0, err.value.decode()
    ref = oracle.nmf_fit_cv(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=1,
                            cd_maxit=20, holdout_fraction=0.1, cv_seed=7, seed=42, mask_zeros=True)
    assert res.iterations == ref.iterations
    assert rel_err(W_out.T, ref.W_T) <= RTOL and rel_err(H_out, ref.H) <= RTOL and rel_err(d, ref.d) <= RTOL
    assert abs(res.test_loss - ref.test_loss) <= 1e-5 * abs(ref.test_loss)
