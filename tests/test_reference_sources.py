"""CPU: the oracle's restatement against the REFERENCE'S OWN hot-path source.

oracle/_ref/libref_hotpath.so is the reference's headers — rng/rng.hpp, primitives/cpu/{nnls_batch, fused_nnls,
cholesky_clip, gram, rhs}.hpp, primitives/primitives.hpp, core/constants.hpp, nmf/masked_nnls.hpp (with core/config.hpp),
nmf/variant_helpers.hpp, features/bounds.hpp, nmf/speckled_cv.hpp, nmf/cv_detail.hpp — compiled unmodified from
/root/reference
(`make -C oracle ref_hotpath`) against a minimal stand-in for the Eigen types they use (Eigen is not in this image).
Everything the reference's source decides — the SplitMix64 generator and hash, the CD solver with its skip rules,
clamps and convergence formula, where L1 / the warm-start correction / the clip / the upper bound sit in the fused
column loops, the Gram wrapper's mirror and 1e-15 — is compared with oracle/nmf_oracle.cpp BIT FOR BIT. The
arithmetic inside Eigen (rankUpdate, gemv, LLT, dot) is implemented in the stand-in with the order the oracle defines
(DESIGN.md §3), so those comparisons pin the call structure, not Eigen's internal rounding.

The library is git-ignored but travels with the repository snapshot; without it (no /root/reference at build time)
the module is skipped."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import random_csc

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_hotpath.so")
pytestmark = pytest.mark.skipif(not os.path.exists(_PATH), reason="oracle/_ref/libref_hotpath.so not built")


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(_PATH)
    lib.ref_splitmix_hash.restype = C.c_uint64
    lib.ref_splitmix_hash.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    lib.ref_is_holdout.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
    lib.ref_loss_cross_term_via_At_f32.restype = C.c_float
    lib.ref_trace_AtA_f32.restype = C.c_float
    lib.ref_masked_loss_f32.restype = C.c_float
    return lib


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def test_constants(ref):
    out = np.zeros(6)
    ref.ref_constants(_p(out, C.c_double))
    assert out[0] == np.float32(1e-15) and out[1] == 1e-15          # tiny_num (core/constants.hpp:42)
    assert out[2] == 1e-8 and out[3] == 100 and out[4] == 1e-15 and out[5] == 5


def test_rng_bit_exact(ref, oracle):
    for seed in (0, 1, 42, 12345, 1234567, 2**63 + 11):
        a = np.zeros(64, np.uint64)
        ref.ref_splitmix_next(C.c_uint64(seed), 64, _p(a, C.c_uint64))
        assert np.array_equal(a, oracle.splitmix_next(seed, 64)), seed
    rng = np.random.default_rng(0)
    for _ in range(2000):
        s, i, j = int(rng.integers(0, 2**62)), int(rng.integers(0, 2**31)), int(rng.integers(0, 2**31))
        assert ref.ref_splitmix_hash(s, i, j) == oracle.splitmix_hash(s, i, j)
        ip = int(rng.integers(1, 50))
        assert bool(ref.ref_is_holdout(s, i, j, ip)) == oracle.is_holdout(s, i, j, ip)
    for seed, r, c in ((42, 7, 13), (12345, 64, 5), (9, 1, 1)):
        f32 = np.zeros((c, r), np.float32)
        ref.ref_fill_uniform_f32(C.c_uint64(seed), _p(f32, C.c_float), r, c)
        assert np.array_equal(f32, oracle.UniformStream(seed).fill_f32(r, c).reshape(c, r)) or \
            np.array_equal(f32.ravel(), oracle.UniformStream(seed).fill_f32(r, c).ravel())
        f64 = np.zeros((c, r), np.float64)
        ref.ref_fill_uniform_f64(C.c_uint64(seed), _p(f64, C.c_double), r, c)
        assert np.array_equal(f64.ravel(), oracle.UniformStream(seed).fill_f64(r, c).ravel())


def test_initialize_factors_bit_exact(ref, oracle):
    for k, m, n, seed in ((7, 30, 20, 42), (64, 100, 10, 12345), (20, 3, 500, 7)):
        W = np.zeros((m, k), np.float32)
        H = np.zeros((n, k), np.float32)
        ref.ref_init_factors_f32(C.c_uint64(seed), k, m, n, _p(W, C.c_float), _p(H, C.c_float))
        oW, oH = oracle.initialize_factors(k, m, n, seed)
        assert np.array_equal(W, oW) and np.array_equal(H, oH)


def _spd(rng, k, dtype, cond_cols=None):
    F = rng.random((cond_cols or 4 * k, k)).astype(dtype)
    return F


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cd_solver_bit_exact(ref, oracle, dtype):
    """cd_nnls_col_fixed: every combination of the switches the reference has, sweeps included."""
    rng = np.random.default_rng(1)
    ct = C.c_float if dtype == np.float32 else C.c_double
    fn = ref.ref_cd_nnls_col_fixed_f32 if dtype == np.float32 else ref.ref_cd_nnls_col_fixed_f64
    for trial in range(60):
        k = int(rng.integers(1, 70))
        F = _spd(rng, k, dtype)
        G = np.ascontiguousarray((F.T @ F).astype(dtype))
        if trial % 7 == 0:
            G[k // 2, k // 2] = 0                                    # a non-positive pivot is skipped (:90)
        b0 = (rng.standard_normal(k) * 3).astype(dtype)
        x0 = np.where(rng.random(k) < 0.4, 0, rng.random(k)).astype(dtype)
        L1 = dtype(0.05 if trial % 3 == 0 else 0.0)
        L2 = dtype(0.1 if trial % 4 == 0 else 0.0)
        nonneg = trial % 5 != 0
        ub = dtype(0.3 if trial % 6 == 0 else 0.0)
        tol = dtype(1e-8 if trial % 2 == 0 else 0.0)
        maxit = int(rng.integers(1, 40))
        b1, x1, b2, x2 = b0.copy(), x0.copy(), b0.copy(), x0.copy()
        s1 = fn(_p(G, ct), _p(b1, ct), _p(x1, ct), k, ct(L1), ct(L2), int(nonneg), maxit, ct(ub), ct(tol))
        s2 = oracle.cd_nnls_col(G, b2, x2, L1=float(L1), L2=float(L2), nonneg=nonneg, maxit=maxit, ub=float(ub),
                                cd_tol=float(tol))
        assert s1 == s2, (trial, s1, s2)
        assert np.array_equal(x1, x2) and np.array_equal(b1, b2), trial


def test_nnls_batch_f64_bit_exact(ref, oracle):
    rng = np.random.default_rng(2)
    for warm in (False, True):
        k, n = 9, 40
        F = rng.random((50, k))
        G = np.ascontiguousarray(F.T @ F)
        B0, X0 = rng.standard_normal((n, k)), rng.random((n, k))
        B1, X1, B2, X2 = B0.copy(), X0.copy(), B0.copy(), X0.copy()
        ref.ref_nnls_batch_f64(_p(G, C.c_double), _p(B1, C.c_double), _p(X1, C.c_double), k, C.c_long(n), 50,
                               C.c_double(1e-8), C.c_double(0.01), C.c_double(0.0), 1, C.c_double(0.0), int(warm))
        oracle.nnls_batch_f64(G, B2, X2, cd_maxit=50, cd_tol=1e-8, L1=0.01, L2=0.0, nonneg=True, ub=0.0, warm_start=warm)
        assert np.array_equal(X1, X2) and np.array_equal(B1, B2)


def test_gram_wrapper_bit_exact(ref, oracle):
    rng = np.random.default_rng(3)
    for k, n in ((6, 100), (20, 333), (64, 50)):
        F = rng.random((n, k)).astype(np.float32)
        G = np.zeros((k, k), np.float32)
        ref.ref_gram_f32(_p(F, C.c_float), k, C.c_long(n), _p(G, C.c_float))
        assert np.array_equal(G, oracle.gram(F))
        assert np.array_equal(G, G.T)


@pytest.mark.parametrize("k", [3, 16, 20, 64])
def test_fused_column_loops_bit_exact(ref, oracle, k):
    """fused_rhs_nnls_sparse (CD; cold and warm) and fused_rhs_cholesky_sparse: L1, non-negativity, upper bound."""
    m, n = 300, 120
    A = random_csc(m, n, 0.06, 40 + k, ragged=True)
    rng = np.random.default_rng(k)
    F = rng.random((m, k)).astype(np.float32)                       # row r = Factor.col(r)
    G = oracle.gram(F)
    X0 = rng.random((n, k)).astype(np.float32)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    for warm in (False, True):
        for L1, nonneg, ub in ((0.0, True, 0.0), (0.02, True, 0.0), (0.0, False, 0.0), (0.01, True, 0.4)):
            X1, X2 = X0.copy(), X0.copy()
            ref.ref_fused_rhs_nnls_sparse_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(m),
                                              C.c_long(n), _p(F, C.c_float), _p(G, C.c_float), _p(X1, C.c_float), k,
                                              30, C.c_float(1e-8), C.c_float(L1), int(nonneg), int(warm), C.c_float(ub))
            oracle.half_step(Ap, Ai, Ax, F, G, X2, solver_mode=0, cd_maxit=30, cd_tol=1e-8, L1=L1, nonneg=nonneg,
                             warm_start=warm, ub_in_solver=ub)
            assert np.array_equal(X1, X2), (k, warm, L1, nonneg, ub)
    for L1, nonneg in ((0.0, True), (0.03, True), (0.0, False)):
        X1, X2 = X0.copy(), X0.copy()
        ref.ref_fused_rhs_cholesky_sparse_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(m),
                                              C.c_long(n), _p(F, C.c_float), _p(G, C.c_float), _p(X1, C.c_float), k, 1,
                                              C.c_float(L1), int(nonneg), C.c_float(0.0))
        oracle.half_step(Ap, Ai, Ax, F, G, X2, solver_mode=1, L1=L1, nonneg=nonneg, warm_start=True)
        assert np.array_equal(X1, X2), (k, L1, nonneg)


def test_loss_terms_match(ref, oracle):
    """trace_AtA and the Gram-trick cross term: the reference accumulates both in fp32 (OpenMP reduction order
    unspecified, sequential here); the oracle defines them in fp64 — agreement to fp32 accumulation error."""
    m, n, k = 400, 150, 12
    A = random_csc(m, n, 0.05, 9)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    tr = ref.ref_trace_AtA_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(m), C.c_long(n))
    assert abs(tr - oracle.trace_AtA(Ax)) <= 1e-5 * abs(tr)
    Tp, Ti, Tx = oracle.transpose_csc(Ap, Ai, Ax, m, n)
    rng = np.random.default_rng(4)
    W, H, d = rng.random((m, k)).astype(np.float32), rng.random((n, k)).astype(np.float32), rng.random(k).astype(np.float32)
    c1 = ref.ref_loss_cross_term_via_At_f32(_p(Tp, C.c_int), _p(Ti, C.c_int), _p(Tx, C.c_float), C.c_long(n), C.c_long(m),
                                            _p(W, C.c_float), _p(H, C.c_float), _p(d, C.c_float), k)
    c2 = oracle.loss_cross_term(Tp, Ti, Tx, W, H, d)
    assert abs(c1 - c2) <= 2e-5 * abs(c2), (c1, c2)


def _mask_pattern(rng, m, n, frac):
    import scipy.sparse as sp
    M = sp.random(m, n, density=frac, format="csc", random_state=rng, dtype=np.float32)
    M.data[:] = 1
    M.sort_indices()
    return M


@pytest.mark.parametrize("k,solver", [(5, 0), (5, 1), (20, 0), (32, 1)])
def test_masked_path_bit_exact(ref, oracle, k, solver):
    """nmf/masked_nnls.hpp (explicit user mask): masked_nnls_h and masked_nnls_w — the per-column Gram downdate over the
    masked rows, L1 on b and L2 on the local diagonal, cold / warm start, CD or the per-column LLT — bit for bit;
    masked_loss to fp32 accumulation error (the reference sums in fp32, the oracle in fp64)."""
    m, n = 160, 90
    rng = np.random.default_rng(100 + k)
    A = random_csc(m, n, 0.12, 70 + k, ragged=True)
    M = _mask_pattern(rng, m, n, 0.05)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    Mp, Mi = M.indptr.astype(np.int32), M.indices.astype(np.int32)
    W = rng.random((m, k)).astype(np.float32)
    H0 = rng.random((n, k)).astype(np.float32)
    G = oracle.gram(W)
    for warm in (False, True):
        H1, H2 = H0.copy(), H0.copy()
        ref.ref_masked_nnls_h_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(m), C.c_long(n),
                                  _p(W, C.c_float), _p(G, C.c_float), _p(H1, C.c_float), k, _p(Mp, C.c_int), _p(Mi, C.c_int),
                                  C.c_float(0.01), C.c_float(0.02), 1, 25, C.c_float(1e-8), solver, int(warm))
        oracle.masked_nnls(Ap, Ai, Ax, m, W, G, H2, Mp, Mi, L1=0.01, L2=0.02, nonneg=True, cd_maxit=25, cd_tol=1e-8,
                           solver_mode=solver, warm_start=warm)
        assert np.array_equal(H1, H2), (k, solver, warm)
    # W half-step: the same kernel on the transposes
    Tp, Ti, Tx = oracle.transpose_csc(Ap, Ai, Ax, m, n)
    MT = M.T.tocsc()
    MT.sort_indices()
    MTp, MTi = MT.indptr.astype(np.int32), MT.indices.astype(np.int32)
    Gh = oracle.gram(H0)
    W1, W2 = W.copy(), W.copy()
    ref.ref_masked_nnls_w_f32(_p(Tp, C.c_int), _p(Ti, C.c_int), _p(Tx, C.c_float), C.c_long(n), C.c_long(m),
                              _p(H0, C.c_float), _p(Gh, C.c_float), _p(W1, C.c_float), k, _p(MTp, C.c_int), _p(MTi, C.c_int),
                              C.c_float(0.0), C.c_float(0.01), 1, 25, C.c_float(1e-8), solver, 1)
    oracle.masked_nnls(Tp, Ti, Tx, n, H0, Gh, W2, MTp, MTi, L1=0.0, L2=0.01, nonneg=True, cd_maxit=25, cd_tol=1e-8,
                       solver_mode=solver, warm_start=True)
    assert np.array_equal(W1, W2), (k, solver)
    # masked loss through a full masked fit of the oracle (one iteration): compare its loss with the reference's sum
    d = rng.random(k).astype(np.float32)
    WTd = (W * d).astype(np.float32)
    l_ref = ref.ref_masked_loss_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), C.c_long(m), C.c_long(n),
                                    _p(WTd, C.c_float), _p(H0, C.c_float), k, _p(Mp, C.c_int), _p(Mi, C.c_int))
    mask_dense = M.toarray() != 0
    pred = WTd.astype(np.float64) @ H0.astype(np.float64).T
    Ad = A.toarray()
    sel = (Ad != 0) & ~mask_dense
    want = float(((Ad - pred)[sel] ** 2).sum())
    assert abs(l_ref - want) <= 1e-4 * want, (l_ref, want)


def test_extract_scaling_and_bounds_bit_exact(ref, oracle):
    rng = np.random.default_rng(8)
    for k, n in ((6, 200), (20, 1000), (64, 77)):
        X0 = (rng.random((n, k)) * (rng.random((n, k)) < 0.6)).astype(np.float32)
        for norm_type in (0, 1, 2):
            X1, X2 = X0.copy(), X0.copy()
            d1 = np.zeros(k, np.float32)
            ref.ref_extract_scaling_f32(_p(X1, C.c_float), k, C.c_long(n), _p(d1, C.c_float), norm_type)
            d2 = oracle.extract_scaling(X2, norm_type)
            assert np.array_equal(d1, d2) and np.array_equal(X1, X2), (k, norm_type)
        X1 = X0.copy()
        ref.ref_apply_upper_bound_f32(_p(X1, C.c_float), k, C.c_long(n), C.c_float(0.3))
        assert np.array_equal(X1, np.minimum(X0, np.float32(0.3)))


def test_speckled_mask_conventions(ref, oracle):
    """LazySpeckledMask: the seed is truncated to 32 bits with 0 -> 12345, inv_prob = uint64(1 / double(fraction)) — the
    reference passes the FLOAT fraction, so 0.1f gives 9 (11.1 % held out), 0.25f gives 4, 0.2f gives 4 (not 5)."""
    rng = np.random.default_rng(5)
    ii = rng.integers(0, 50000, 4000).astype(np.int32)
    jj = rng.integers(0, 9000, 4000).astype(np.int32)
    for frac, inv in ((0.1, 9), (0.25, 4), (0.2, 4), (0.05, 19), (0.5, 2)):
        for seed in (0, 42, 2**32 + 7, 2**32):
            out = np.zeros(4000, np.int32)
            ref.ref_speckled_mask_f32(50000, 9000, C.c_double(float(np.float32(frac))), C.c_uint64(seed),
                                      _p(ii, C.c_int), _p(jj, C.c_int), 4000, _p(out, C.c_int))
            s32 = seed & 0xFFFFFFFF
            eff = 12345 if s32 == 0 else s32
            want = np.array([oracle.is_holdout(eff, int(i), int(j), inv) for i, j in zip(ii, jj)], dtype=np.int32)
            assert np.array_equal(out, want), (frac, seed)
            assert 0.5 / inv < out.mean() < 1.5 / inv


@pytest.mark.parametrize("mask_zeros", [True, False])
def test_cv_building_blocks_bit_exact(ref, oracle, mask_zeros):
    """nmf/cv_detail.hpp: compute_train_rhs (H side: operand A, mask(row, col)), compute_train_rhs_W (W side: operand
    Aᵀ, mask(row of A = column of Aᵀ, col = inner)) and apply_gram_correction — which entries are held out, in which
    order, what accumulates into b (mask_zeros: only stored entries can be held out; otherwise every cell of the
    column is hashed and a held-out structural zero counts too), and the corrected Gram."""
    from oracle.oracle import lib as olib
    L = olib()
    L.orc_cv_train_rhs_f32.restype = C.c_long
    ref.ref_cv_train_rhs_f32.restype = C.c_long
    m, n, k = 140, 60, 9
    A = random_csc(m, n, 0.15, 31, ragged=True)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    Tp, Ti, Tx = oracle.transpose_csc(Ap, Ai, Ax, m, n)
    rng = np.random.default_rng(12)
    W = rng.random((m, k)).astype(np.float32)
    H = rng.random((n, k)).astype(np.float32)
    frac = float(np.float32(0.2))                       # the reference holds the fraction as a float: inv_prob = 4
    seed, inv_prob = 77, 4
    for transposed, (Cp, Ci, Cx, n_inner, n_cols, F) in ((0, (Ap, Ai, Ax, m, n, W)), (1, (Tp, Ti, Tx, n, m, H))):
        G = oracle.gram(F)
        for col in range(0, n_cols, 7):
            b1, b2 = np.zeros(k, np.float32), np.zeros(k, np.float32)
            t1, t2 = np.zeros(n_inner + 1, np.int32), np.zeros(n_inner + 1, np.int32)
            c1 = ref.ref_cv_train_rhs_f32(_p(Cp, C.c_int), _p(Ci, C.c_int), _p(Cx, C.c_float), C.c_long(n_inner),
                                          C.c_long(n_cols), C.c_long(col), _p(F, C.c_float), k, transposed,
                                          int(mask_zeros), C.c_double(frac), C.c_uint64(seed), _p(b1, C.c_float), _p(t1, C.c_int))
            c2 = L.orc_cv_train_rhs_f32(_p(Cp, C.c_int), _p(Ci, C.c_int), _p(Cx, C.c_float), C.c_long(n_inner),
                                        C.c_long(col), _p(F, C.c_float), k, transposed, int(mask_zeros),
                                        C.c_uint64(seed), C.c_uint64(inv_prob), _p(b2, C.c_float), _p(t2, C.c_int))
            assert c1 == c2 and np.array_equal(t1[:c1], t2[:c2]), (transposed, col, c1, c2)
            assert np.array_equal(b1, b2), (transposed, col)
            G1, G2 = np.zeros((k, k), np.float32), np.zeros((k, k), np.float32)
            ref.ref_cv_gram_correction_f32(_p(G, C.c_float), _p(F, C.c_float), C.c_long(n_inner), _p(t1, C.c_int),
                                           C.c_long(c1), k, _p(G1, C.c_float))
            L.orc_cv_gram_correction_f32(_p(G, C.c_float), _p(F, C.c_float), _p(t2, C.c_int), C.c_long(c2), k, _p(G2, C.c_float))
            assert np.array_equal(G1, G2), (transposed, col)
        if not mask_zeros:
            assert c1 > 0                                # with every cell hashed a column of 60+ cells has hold-outs


@pytest.mark.parametrize("k", [5, 20, 64])
def test_predict_nnls_pipeline_bit_exact(ref, oracle, k):
    """predict() / nnls(): the bodies of c_nnls / Rcpp_predict (src/RcppFunctions_utils.cpp:314-366,
    23-53 — gram<CPU,double>, + tiny_num, + L2, rhs<CPU,double>, nnls_batch<CPU,double> or the warm-started CD loop)
    assembled from the reference's own primitives, against the oracle's project_f64 — the checker of the GPU entry
    rcppml_gpu_nnls_double."""
    if not hasattr(ref, "ref_c_nnls_sparse_f64"):
        pytest.skip("oracle/_ref/libref_hotpath.so predates ref_c_nnls_sparse_f64 (rebuild: make -C oracle ref_hotpath)")
    m, n = 300, 120
    A = random_csc(m, n, 0.08, 90 + k, counts=True, ragged=True)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    rng = np.random.default_rng(k)
    w = rng.random((m, k))                                                 # (m, k) C-order == k x m column-major
    cases = [dict(), dict(L1=0.05, L2=0.1), dict(upper_bound=0.3), dict(nonneg=False, L2=0.01), dict(cd_maxit=3),
             dict(warm_start=rng.random((n, k)) * 0.1, cd_maxit=7), dict(warm_start=rng.random((n, k)), L1=0.02, upper_bound=0.5)]
    for kw in cases:
        got = oracle.project_f64(Ap, Ai, Ax, m, n, w, **kw)               # (n, k)
        ws = kw.get("warm_start")
        h = np.zeros((n, k)) if ws is None else np.array(ws, np.float64, order="C")
        ref.ref_c_nnls_sparse_f64(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_double), C.c_long(m), C.c_long(n),
                                  _p(w, C.c_double), k, _p(h, C.c_double), kw.get("cd_maxit", 100),
                                  C.c_double(kw.get("cd_tol", 1e-8)), C.c_double(kw.get("L1", 0.0)),
                                  C.c_double(kw.get("L2", 0.0)), C.c_double(kw.get("upper_bound", 0.0)),
                                  int(kw.get("nonneg", True)), int(ws is not None))
        assert np.array_equal(got, h), (k, {a: b for a, b in kw.items() if a != "warm_start"}, float(np.abs(got - h).max()))
