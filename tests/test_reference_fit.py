"""CPU: the oracle against THE REFERENCE'S OWN ALS LOOPS (nmf_fit and, at the end of the file, nmf_fit_cv).

oracle/_ref/libref_fit.so is the reference's nmf/fit_cpu.hpp — nmf_fit<CPU, float, SparseMatrix<float>>, 1855 lines,
with every header it pulls in — compiled unmodified from /root/reference (`make -C oracle ref_hotpath`) against the
minimal Eigen stand-in of oracle/ref_hotpath/shim (Eigen is not in this image); the headers of features outside the
hot path (SVD initialisation, IRLS losses, graph / L21 / angular regularisers) are shadowed by declarations that throw
if called (oracle/ref_hotpath/out_of_path). So the ORCHESTRATION of a fit — initialisation, transpose, which Gram
feeds which half-step, where L1 / L2 / bounds / scaling sit, the fused-path predicate and its first-iteration quirk, the
explicit-mask branch, the Gram-trick loss, patience, result packing, sorting — is the reference's own code here, and
the oracle (oracle/nmf_oracle.cpp) must reproduce its W, d, H BIT FOR BIT.

Only the arithmetic inside Eigen is a definition shared by both sides. The loss is compared to 1e-5:
the reference accumulates tr(AᵀA) and the cross term in fp32, the oracle (and the GPU) in fp64."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import random_csc

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_fit.so")
pytestmark = pytest.mark.skipif(not os.path.exists(_PATH), reason="oracle/_ref/libref_fit.so not built")


class P(C.Structure):
    _fields_ = [("k", C.c_int), ("max_iter", C.c_int), ("tol", C.c_float), ("L1_W", C.c_float), ("L1_H", C.c_float),
                ("L2_W", C.c_float), ("L2_H", C.c_float), ("ub_W", C.c_float), ("ub_H", C.c_float),
                ("nonneg_W", C.c_int), ("nonneg_H", C.c_int), ("cd_maxit", C.c_int), ("cd_tol", C.c_float),
                ("norm_type", C.c_int), ("solver_mode", C.c_int), ("patience", C.c_int), ("threads", C.c_int),
                ("sort_model", C.c_int), ("seed", C.c_uint)]


class R(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("train_loss", C.c_float), ("final_tol", C.c_float),
                ("n_loss", C.c_int)]


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


@pytest.fixture(scope="module")
def reffit():
    return C.CDLL(_PATH)


def _ref_fit(lib, A, k, W0, H0, *, max_iter, tol=0.0, L1=(0.0, 0.0), L2=(0.0, 0.0), upper_bound=(0.0, 0.0),
             nonneg=(True, True), cd_maxit=100, cd_tol=1e-8, norm_type=0, solver_mode=0, patience=5, sort_model=False,
             mask=None, seed=42):
    m, n = A.shape
    q = P(k=k, max_iter=max_iter, tol=tol, L1_W=L1[0], L1_H=L1[1], L2_W=L2[0], L2_H=L2[1], ub_W=upper_bound[0],
          ub_H=upper_bound[1], nonneg_W=int(nonneg[0]), nonneg_H=int(nonneg[1]), cd_maxit=cd_maxit, cd_tol=cd_tol,
          norm_type=norm_type, solver_mode=solver_mode, patience=patience, threads=1, sort_model=int(sort_model), seed=seed)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    W_in = np.ascontiguousarray(W0.T)                               # m x k column-major
    H_in = None if H0 is None else np.ascontiguousarray(H0)        # k x n column-major
    Mp = Mi = None
    if mask is not None:
        Mp, Mi = mask[0].astype(np.int32), mask[1].astype(np.int32)
    W_out, H_out, d = np.zeros((k, m), np.float32), np.zeros((n, k), np.float32), np.zeros(k, np.float32)
    hist = np.zeros(max_iter, np.float32)
    res, err = R(), C.create_string_buffer(300)
    rc = lib.reffit_nmf_sparse_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), m, n, C.byref(q),
                                   _p(W_in, C.c_float), _p(H_in, C.c_float), _p(Mp, C.c_int), _p(Mi, C.c_int),
                                   _p(W_out, C.c_float), _p(H_out, C.c_float), _p(d, C.c_float), _p(hist, C.c_float),
                                   C.byref(res), err, 300)
    assert rc == 0, err.value.decode()
    return W_out.T.copy(), H_out, d, hist[:res.n_loss].copy(), res


CASES = [
    ("cd_k8", 300, 200, 0.10, 8, dict(solver_mode=0)),
    ("chol_k8", 300, 200, 0.10, 8, dict(solver_mode=1)),
    ("cd_k20_L1", 400, 260, 0.06, 20, dict(solver_mode=0, L1=(0.01, 0.01))),
    ("chol_k20_L1L2", 400, 260, 0.06, 20, dict(solver_mode=1, L1=(0.01, 0.02), L2=(0.02, 0.01))),
    ("cd_k32_L2", 500, 300, 0.05, 32, dict(solver_mode=0, L2=(0.01, 0.0))),
    ("chol_k64", 600, 350, 0.05, 64, dict(solver_mode=1)),
    ("cd_k5_ub_l2norm", 200, 150, 0.12, 5, dict(solver_mode=0, upper_bound=(0.05, 0.08), norm_type=1)),
    ("chol_k6_ub_nonorm", 200, 150, 0.12, 6, dict(solver_mode=1, upper_bound=(0.3, 0.2), norm_type=2)),
    ("cd_k7_seminmf", 200, 150, 0.12, 7, dict(solver_mode=0, nonneg=(True, False))),
    ("chol_k7_seminmf_sorted", 200, 150, 0.12, 7, dict(solver_mode=1, nonneg=(False, True), sort_model=True)),
    ("cd_k12_few_sweeps", 300, 200, 0.08, 12, dict(solver_mode=0, cd_maxit=3)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_the_reference_fit_bit_for_bit(reffit, oracle, case):
    name, m, n, dens, k, kw = case
    iters = 7
    A = random_csc(m, n, dens, 17, counts=("L1" in name), ragged=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    W, H, d, hist, res = _ref_fit(reffit, A, k, W0, H0, max_iter=iters, **kw)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, threads=1, **kw)
    assert res.iterations == ref.iterations == iters
    assert np.array_equal(W, ref.W_T), (name, float(np.abs(W - ref.W_T).max()))
    assert np.array_equal(H, ref.H), (name, float(np.abs(H - ref.H).max()))
    assert np.array_equal(d, ref.d), name
    assert len(hist) == iters and np.allclose(hist, ref.loss_history, rtol=1e-5, atol=0), (hist, ref.loss_history)


def test_reference_fit_with_its_own_h_initialisation(reffit, oracle):
    """H_init = nullptr: the reference draws H from a fresh SplitMix64(seed) (fit_cpu.hpp:203-206)."""
    m, n, k, iters = 250, 180, 9, 5
    A = random_csc(m, n, 0.1, 4)
    W0, _ = oracle.initialize_factors(k, m, n, 42)
    H0 = oracle.init_h_given_w(k, n, 42)
    for solver in (0, 1):
        W, H, d, hist, res = _ref_fit(reffit, A, k, W0, None, max_iter=iters, solver_mode=solver, seed=42)
        ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver, threads=1)
        assert np.array_equal(W, ref.W_T) and np.array_equal(H, ref.H) and np.array_equal(d, ref.d)


@pytest.mark.parametrize("solver", [0, 1])
def test_reference_fit_with_an_explicit_mask(reffit, oracle, solver):
    """config.mask: masked_nnls_h / masked_nnls_w in both half-steps and the explicit masked loss (fit_cpu.hpp:560-564,
    :799-810, :1686-1691)."""
    import scipy.sparse as sp
    m, n, k, iters = 220, 160, 6, 5
    A = random_csc(m, n, 0.12, 8, ragged=True)
    M = sp.random(m, n, density=0.05, format="csc", random_state=np.random.default_rng(3), dtype=np.float32)
    M.sort_indices()
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    kw = dict(solver_mode=solver, L1=(0.01, 0.0), L2=(0.0, 0.01))
    W, H, d, hist, res = _ref_fit(reffit, A, k, W0, H0, max_iter=iters, mask=(M.indptr, M.indices), **kw)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, threads=1,
                         mask=(M.indptr, M.indices), **kw)
    assert np.array_equal(W, ref.W_T) and np.array_equal(H, ref.H) and np.array_equal(d, ref.d)
    assert np.allclose(hist, ref.loss_history, rtol=1e-5, atol=0)


def test_reference_fit_convergence_and_patience(reffit, oracle):
    """tol > 0: the reference stops after `patience` consecutive relative changes below tol (fit_cpu.hpp:1769-1809);
    same iteration count, same flag, same factors."""
    m, n, k = 200, 140, 5
    A = random_csc(m, n, 0.15, 21)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    for patience in (1, 3):
        W, H, d, hist, res = _ref_fit(reffit, A, k, W0, H0, max_iter=200, tol=2e-4, patience=patience, solver_mode=1)
        ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=200, tol=2e-4, patience=patience,
                             solver_mode=1, threads=1)
        assert res.converged == 1 and ref.converged and res.iterations == ref.iterations < 200
        assert np.array_equal(W, ref.W_T) and np.array_equal(H, ref.H) and np.array_equal(d, ref.d)
        # the relative change is a difference of nearly equal losses that the two sides accumulate in fp32 / fp64
        assert abs(res.final_tol - ref.final_tol) <= 2e-2 * abs(ref.final_tol) + 1e-9


# ---- the cross-validation orchestrator: nmf/fit_cv.hpp, nmf_fit_cv<CPU, float, SparseMatrix<float>> ----------------
class CP(C.Structure):
    _fields_ = [("holdout_fraction", C.c_float), ("cv_seed", C.c_uint), ("mask_zeros", C.c_int), ("cv_patience", C.c_int)]


class CR(C.Structure):
    _fields_ = [("test_loss", C.c_float), ("best_test_loss", C.c_float), ("best_iter", C.c_int), ("n_test_hist", C.c_int)]


CV_CASES = [(7, 1, True, dict(L1=(0.01, 0.0), L2=(0.0, 0.01))), (7, 0, True, {}), (5, 1, False, {}),
            (12, 0, False, dict(L1=(0.0, 0.01))), (20, 1, True, dict(upper_bound=(0.5, 0.5))), (6, 0, True, dict(norm_type=1))]


@pytest.mark.parametrize("k,solver,mask_zeros,kw", CV_CASES, ids=[f"k{c[0]}_s{c[1]}_mz{int(c[2])}" for c in CV_CASES])
def test_oracle_reproduces_the_reference_cv_fit_bit_for_bit(reffit, oracle, k, solver, mask_zeros, kw):
    """Speckled-mask cross-validation: W, d, H bit for bit, same best iteration; train / test losses to 1e-5 (the
    reference sums them in fp32, the oracle in fp64)."""
    m, n, iters = 260, 180, 6
    A = random_csc(m, n, 0.12, 3 + k, ragged=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    L1, L2, ub = kw.get("L1", (0.0, 0.0)), kw.get("L2", (0.0, 0.0)), kw.get("upper_bound", (0.0, 0.0))
    q = P(k=k, max_iter=iters, tol=0.0, L1_W=L1[0], L1_H=L1[1], L2_W=L2[0], L2_H=L2[1], ub_W=ub[0], ub_H=ub[1],
          nonneg_W=1, nonneg_H=1, cd_maxit=15, cd_tol=1e-8, norm_type=kw.get("norm_type", 0), solver_mode=solver,
          patience=5, threads=1, sort_model=0, seed=42)
    cq = CP(holdout_fraction=0.1, cv_seed=7, mask_zeros=int(mask_zeros), cv_patience=5)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    W_in, H_in = np.ascontiguousarray(W0.T), np.ascontiguousarray(H0)
    W_out, H_out, d = np.zeros((k, m), np.float32), np.zeros((n, k), np.float32), np.zeros(k, np.float32)
    tr, te = np.zeros(iters, np.float32), np.zeros(iters, np.float32)
    res, cres, err = R(), CR(), C.create_string_buffer(300)
    rc = reffit.reffit_nmf_cv_sparse_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), m, n, C.byref(q), C.byref(cq),
                                         _p(W_in, C.c_float), _p(H_in, C.c_float), _p(W_out, C.c_float),
                                         _p(H_out, C.c_float), _p(d, C.c_float), _p(tr, C.c_float), _p(te, C.c_float),
                                         C.byref(res), C.byref(cres), err, 300)
    assert rc == 0, err.value.decode()
    ref = oracle.nmf_fit_cv(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver,
                            cd_maxit=15, holdout_fraction=0.1, cv_seed=7, seed=42, mask_zeros=mask_zeros, threads=1, **kw)
    assert res.iterations == ref.iterations and cres.best_iter == ref.best_iter
    assert np.array_equal(W_out.T, ref.W_T), float(np.abs(W_out.T - ref.W_T).max())
    assert np.array_equal(H_out, ref.H) and np.array_equal(d, ref.d)
    assert np.allclose(te[:cres.n_test_hist], ref.test_history, rtol=1e-5, atol=0)
    assert np.allclose(tr[:res.n_loss], ref.train_history, rtol=1e-5, atol=0)
    assert abs(cres.best_test_loss - ref.best_test_loss) <= 1e-5 * abs(ref.best_test_loss)


# ---- the GPU engine against the reference's own fit, directly ----------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if "sorted" not in c[0]], ids=[c[0] for c in CASES if "sorted" not in c[0]])
def test_gpu_engine_matches_the_reference_fit(reffit, oracle, case):
    """north_star: "W, d, H match the reference CPU path within 1e-5 relative fp32 tolerance" — here the reference CPU
    path is the reference's own nmf_fit (libref_fit.so), not the restatement. (sort_model is not on the GPU wire.)"""
    import rcppml_b200 as rb
    from helpers import RTOL, rel_err, zero_pattern_equal
    name, m, n, dens, k, kw = case
    iters = 7
    A = random_csc(m, n, dens, 17, counts=("L1" in name), ragged=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    W, H, d, hist, res = _ref_fit(reffit, A, k, W0, H0, max_iter=iters, **kw)
    eng = rb.Engine(0)
    try:
        eng.set_matrix(m, n, A.indptr, A.indices, A.data)
        eng.set_factors(W0, H0)
        out = eng.fit(rb.make_config(k, max_iter=iters, tol=0.0, **kw))
        Wg, Hg, dg = eng.get_factors()
        lg = eng.loss_history(iters)
    finally:
        eng.close()
    assert out.status == 0 and out.iterations == iters
    errs = dict(W=rel_err(Wg, W), H=rel_err(Hg, H), d=rel_err(dg, d), loss=rel_err(lg, hist))
    assert max(errs.values()) <= RTOL, (name, errs)
    assert zero_pattern_equal(Wg, W) and zero_pattern_equal(Hg, H)
