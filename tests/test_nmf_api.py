"""rcppml_b200.nmf(): the host mirror of the reference's R front end (R/nmf_thin.R `nmf`) over the GPU bridge.

CPU part: R's RNG (set.seed + runif known answers), the bridge's H stream, argument validation with the reference's
messages, refusal of everything outside the hot path, no CPU fallback.
GPU part reads like the reference's own tests (tests/testthat/test_nmf.R, test_norm.R, test_upper_bound.R,
test_masking.R, test_unified_backend.R) plus oracle parity of the whole nmf() call: same W (R's runif), same H
(bridge stream), same factors to 1e-5."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import RTOL, random_csc, rel_err, zero_pattern_equal


# ------------------------------------------------------------------------------------------------ CPU ----
def test_r_mersenne_twister_known_answers():
    """`set.seed(s); runif(3)` as every R session prints them (R >= 3.6 default RNG kind)."""
    from rcppml_b200.nmf import RRandom
    assert np.allclose(RRandom(42).runif(3), [0.9148060435, 0.9370754133, 0.2861395348], atol=5e-9)
    assert np.allclose(RRandom(1).runif(3), [0.2655086631, 0.3721238966, 0.5728533633], atol=5e-9)
    assert np.allclose(RRandom(123).runif(3), [0.2875775201, 0.7883051354, 0.4089769218], atol=5e-9)
    u = RRandom(7).runif(100000)
    assert u.min() > 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    assert np.array_equal(RRandom(-3).runif(4), RRandom(2 ** 32 - 3).runif(4))       # Int32 wrap of the seed


def test_bridge_h_stream_matches_the_oracle_rng(oracle):
    """gpu/bridge_nmf.hpp:226-229 — SplitMix64(seed + 0x9E3779B9u) drawn as doubles."""
    from rcppml_b200.nmf import bridge_h_init, bridge_w_init, splitmix_fill_uniform_f64
    for seed in (42, 0, 123456789, 2 ** 31 - 1):
        k, n = 5, 37
        ref = oracle.UniformStream((seed + 0x9E3779B9) & 0xFFFFFFFF).fill_f64(k, n)
        assert np.array_equal(bridge_h_init(seed, k, n), ref.reshape(n, k))
    assert np.array_equal(splitmix_fill_uniform_f64(0, 11), oracle.UniformStream(12345).fill_f64(11, 1).reshape(-1))
    assert np.array_equal(bridge_w_init(42, 4, 9), oracle.UniformStream(42).fill_f64(4, 9).reshape(9, 4))


def test_nmf_argument_validation_uses_the_reference_messages():
    import rcppml_b200 as rb
    A = random_csc(30, 20, 0.3, 1)
    cases = [
        (dict(L1=(0.1, 0.2, 0.3)), ValueError, "'L1' must be length 1 or 2 for c(w, h)"),
        (dict(L1=1.0), ValueError, "L1 penalties must be strictly in the range [0,1)"),
        (dict(L2=-1), ValueError, "'L2' values must be non-negative"),
        (dict(upper_bound=(-1, 0)), ValueError, "'upper_bound' values must be non-negative"),
        (dict(test_fraction=1.0), ValueError, "'test_fraction' must be in the range [0, 1)"),
        (dict(test_fraction="a"), ValueError, "'test_fraction' must be a single numeric value"),
        (dict(nonneg=(1, 0)), ValueError, "'nonneg' must be logical"),
        (dict(nonneg=(True, True, False)), ValueError, "'nonneg' must be length 1 or 2 with no NA values"),
        (dict(sort_model="yes"), ValueError, "'sort_model' must be a single logical value"),
        (dict(mask="ones"), ValueError, "'mask' must be NULL, 'zeros', 'NA', a matrix, or list"),
        (dict(mask=["ones", A]), ValueError, "'mask' list must be list(\"zeros\", <matrix>)"),
        (dict(seed="kmeans"), ValueError, "Unknown seed string 'kmeans'"),
        (dict(seed=np.ones((30, 4))), ValueError, "Rank mismatch: k=3 specified but custom initialization has rank 4."),
        (dict(bogus=1, other=2), TypeError, "Unknown parameter(s) passed to nmf(): 'bogus', 'other'. See ?nmf"),
        (dict(resource="tpu"), ValueError, "'resource' must be \"auto\", \"cpu\", or \"gpu\""),
        (dict(projective=True, symmetric=True), ValueError, "'projective' and 'symmetric' cannot both be TRUE"),
        (dict(zi="row"), ValueError, "zi != 'none' requires loss='gp' or loss='nb'."),
        (dict(solver="qr"), ValueError, "'solver' should be one of"),
        (dict(w_init=np.ones((7, 7))), ValueError, "w_init dimensions incompatible with data and k"),
        (dict(seed=[1, 2, 3], test_fraction=0.1), ValueError, "Multiple initializations are not compatible with cross-validation"),
        (dict(seed=[np.ones((30, 3)), np.ones(3)]), ValueError, "Each element of seed list must be a matrix"),
        (dict(seed=[np.ones((30, 3)), np.ones((30, 4))]), ValueError, "Rank mismatch: k=3 specified but custom initialization has rank 4."),
    ]
    for kw, exc, msg in cases:
        with pytest.raises(exc) as e:
            rb.nmf(A, 3, **kw)
        assert msg in str(e.value), (kw, str(e.value))


def test_nmf_refuses_what_is_outside_the_hot_path():
    import rcppml_b200 as rb
    A = random_csc(30, 20, 0.3, 1)
    for kw in (dict(loss="gp"), dict(robust=True), dict(projective=True), dict(symmetric=True), dict(L21=0.1),
               dict(angular=(0.0, 0.2)), dict(graph_W=sp.identity(30)), dict(target_H=np.ones((3, 20))),
               dict(seed="lanczos"), dict(resource="cpu"), dict(streaming=True)):
        with pytest.raises(NotImplementedError):
            rb.nmf(A, 3, **kw)
    with pytest.raises(NotImplementedError):
        rb.nmf(A, [2, 3, 4])
    with pytest.raises(NotImplementedError):
        rb.nmf(A, "auto")
    with pytest.raises(NotImplementedError):
        rb.nmf(A.toarray(), 3)
    with pytest.raises(NotImplementedError):
        rb.nmf("matrix.spz", 3)


def test_nmf_has_no_cpu_fallback():
    """Where the reference's gateway falls back to its CPU loop (nmf/fit.hpp:118-127) this library raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    import rcppml_b200 as rb
    from rcppml_b200._lib import NativeLibraryError
    with pytest.raises(NativeLibraryError):
        rb.nmf(random_csc(30, 20, 0.3, 1), 3, seed=1)


def test_sort_by_d_is_result_sort():
    from rcppml_b200.nmf import sort_by_d
    w, d, h = np.arange(12.0).reshape(4, 3), np.array([0.2, 0.7, 0.1]), np.arange(15.0).reshape(3, 5)
    ws, ds, hs = sort_by_d(w, d, h)
    assert np.array_equal(ds, [0.7, 0.2, 0.1])
    assert np.array_equal(ws, w[:, [1, 0, 2]]) and np.array_equal(hs, h[[1, 0, 2], :])


def test_nnls_argument_validation_uses_the_reference_messages():
    """R/solve.R:89-96, 211-213, 282-298."""
    import rcppml_b200 as rb
    A = random_csc(30, 20, 0.3, 1)
    w, h = np.ones((30, 4)), np.ones((4, 20))
    cases = [
        (dict(A=A), ValueError, "Either 'w' or 'h' must be provided (not both NULL)"),
        (dict(w=w, h=h, A=A), ValueError, "Exactly one of 'w' or 'h' must be NULL (cannot provide both)"),
        (dict(w=w, A=A, L1=1.0), ValueError, "L1 penalty must be in range [0, 1)"),
        (dict(w=w, A=A, L2=(-1, 0)), ValueError, "L2 penalty must be >= 0"),
        (dict(w=w, A=A, upper_bound=-1), ValueError, "upper_bound must be >= 0"),
        (dict(w=np.ones((7, 4)), A=A), ValueError, "Incompatible dimensions: nrow(w) = 7 but nrow(A) = 30"),
        (dict(h=np.ones((4, 7)), A=A), ValueError, "Incompatible dimensions: ncol(h) = 7 but ncol(A) = 20"),
        (dict(w=w, A=A, warm_start=np.ones((3, 20))), ValueError, "warm_start dimensions (3 x 20) don't match expected "
                                                                 "output dimensions (4 x 20)"),
        (dict(w=w, A=A, bogus=1), TypeError, "Unknown parameter(s) passed to nnls(): 'bogus'"),
        (dict(w=w, A=A, loss="gp"), NotImplementedError, "nnls(loss"),
        (dict(w=w, A=A, L21=0.1), NotImplementedError, "nnls(loss"),
        (dict(h=h, A=A, warm_start=np.ones((30, 4))), NotImplementedError, "warm_start with h"),
    ]
    for kw, exc, msg in cases:
        with pytest.raises(exc) as e:
            rb.nnls(**kw)
        assert msg in str(e.value), (kw, str(e.value))


# ------------------------------------------------------------------------------------------------ GPU ----
def _oracle_twin(oracle, A, k, seed, **kw):
    """What the reference's CPU loop computes from the inputs nmf() hands to the bridge."""
    from rcppml_b200.nmf import RRandom, bridge_h_init
    m, n = A.shape
    W0 = RRandom(seed).runif(m * k).reshape(k, m).T.astype(np.float32)
    H0 = bridge_h_init(seed, k, n).astype(np.float32)
    return W0, H0, lambda **more: oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, **kw, **more)


@pytest.mark.gpu
@pytest.mark.parametrize("k,kw,okw", [
    (8, dict(), dict(solver_mode=0)),
    (20, dict(L1=0.01), dict(solver_mode=0, L1=(0.01, 0.01))),
    (40, dict(L2=(0.01, 0.0)), dict(solver_mode=1, L2=(0.01, 0.0))),                       # auto -> cholesky above 32
    (6, dict(solver="cholesky", upper_bound=0.3, norm="L2"), dict(solver_mode=1, upper_bound=(0.3, 0.3), norm_type=1)),
    (7, dict(nonneg=(True, False), norm="none"), dict(solver_mode=0, nonneg=(True, False), norm_type=2)),
])
def test_nmf_matches_the_oracle(oracle, k, kw, okw):
    import rcppml_b200 as rb
    A = random_csc(300, 180, 0.1, 11 + k, ragged=True)
    model = rb.nmf(A, k, seed=123, maxit=8, tol=0.0, **kw)
    _, _, fit = _oracle_twin(oracle, A, k, 123, max_iter=8, tol=0.0, **okw)
    ref = fit(sort_model=True)
    assert model.misc["iter"] == 8 and model.w.shape == (300, k) and model.h.shape == (k, 180)
    assert rel_err(model.w, ref.W_T) <= RTOL and rel_err(model.h, ref.H.T) <= RTOL and rel_err(model.d, ref.d) <= RTOL
    assert zero_pattern_equal(model.w, ref.W_T) and zero_pattern_equal(model.h, ref.H.T)
    assert abs(model.misc["loss"] - ref.train_loss) <= 1e-5 * abs(ref.train_loss)
    assert np.all(np.diff(model.d) <= 0)                                                    # sort_model = TRUE
    unsorted = rb.nmf(A, k, seed=123, maxit=8, tol=0.0, sort_model=False, **kw)
    raw = fit(sort_model=False)
    assert rel_err(unsorted.w, raw.W_T) <= RTOL and rel_err(unsorted.d, raw.d) <= RTOL


@pytest.mark.gpu
def test_nmf_seed_forms_and_reproducibility():
    """test_nmf.R: same seed -> same model; a W matrix as seed (either orientation) and w_init are honoured."""
    import rcppml_b200 as rb
    from rcppml_b200.nmf import RRandom, bridge_h_init
    A = random_csc(200, 120, 0.1, 3)
    a = rb.nmf(A, 5, seed=42, maxit=5, tol=0.0)
    b = rb.nmf(A, 5, seed=42, maxit=5, tol=0.0)
    assert np.array_equal(a.w, b.w) and np.array_equal(a.h, b.h) and np.array_equal(a.d, b.d)
    W0 = RRandom(42).runif(200 * 5).reshape(5, 200).T
    assert np.array_equal(a.misc["w_init"], W0)
    kw = dict(maxit=5, tol=0.0, h_init=bridge_h_init(42, 5, 120).T)
    c = rb.nmf(A, 5, seed=W0, **kw)
    e = rb.nmf(A, 5, seed=W0.T.copy(), **kw)
    f = rb.nmf(A, 5, seed=7, w_init=W0, **kw)
    for other in (c, e, f):
        assert np.array_equal(a.w, other.w) and np.array_equal(a.h, other.h)
    g = rb.nmf(A, 5, maxit=3, tol=0.0)                                                     # seed = NULL
    assert g.w.shape == (200, 5) and np.isfinite(g.misc["loss"]) and 0 <= g.misc["seed_int"] < 2 ** 31


@pytest.mark.gpu
def test_nmf_invariants_of_the_reference_tests():
    """test_norm.R:35-52 (L1: colSums(w) = 1 and rowSums(h) = 1; L2: unit norms; none: d = 1), test_upper_bound.R:24-41,
    test_nmf.R (non-negativity, tol stops early), test_unified_backend.R:241-284 (loss history non-increasing)."""
    import rcppml_b200 as rb
    A = random_csc(250, 150, 0.12, 9, counts=True)
    m1 = rb.nmf(A, 6, seed=1, maxit=30, tol=1e-5)
    assert np.allclose(m1.w.sum(axis=0), 1.0, atol=1e-5) and np.allclose(m1.h.sum(axis=1), 1.0, atol=1e-5)
    assert m1.w.min() >= 0 and m1.h.min() >= 0 and m1.d.min() > 0
    m2 = rb.nmf(A, 6, seed=1, maxit=10, tol=0.0, norm="L2")
    assert np.allclose(np.linalg.norm(m2.w, axis=0), 1.0, atol=1e-5)
    assert np.allclose(np.linalg.norm(m2.h, axis=1), 1.0, atol=1e-5)
    m3 = rb.nmf(A, 6, seed=1, maxit=10, tol=0.0, norm="none")
    assert np.allclose(m3.d, 1.0, atol=1e-6)
    m4 = rb.nmf(A, 6, seed=1, maxit=10, tol=0.0, upper_bound=(0.02, 0.05), norm="none")
    assert m4.w.max() <= 0.02 + 1e-7 and m4.h.max() <= 0.05 + 1e-7
    loose = rb.nmf(A, 6, seed=1, maxit=200, tol=1e-2)
    assert loose.misc["iter"] < 200 and loose.misc["converged"]
    r3, r10 = rb.nmf(A, 6, seed=1, maxit=3, tol=0.0), rb.nmf(A, 6, seed=1, maxit=10, tol=0.0)
    assert r10.misc["loss"] <= r3.misc["loss"]
    mse = m1.evaluate(A)
    dense = A.toarray().astype(np.float64)
    assert abs(mse - np.mean((dense - m1.reconstruct()) ** 2)) <= 1e-6 * max(mse, 1e-12)
    proj = m1.predict(A)
    assert proj.h.shape == (6, 150) and proj.h.min() >= 0 and proj.misc == dict(projected=True) and proj.w is m1.w
    mz = np.mean((dense - m1.reconstruct())[dense != 0] ** 2)
    assert abs(m1.evaluate(A, mask="zeros") - mz) <= 1e-6 * mz


@pytest.mark.gpu
def test_nmf_explicit_mask_and_mask_zeros(oracle):
    """test_masking.R: an explicit mask changes the fit (masked entries are not fitted); mask = "zeros" does not
    change a non-CV fit."""
    import rcppml_b200 as rb
    A = random_csc(220, 140, 0.12, 21, ragged=True)
    rng = np.random.default_rng(5)
    M = sp.random(220, 140, density=0.05, format="csc", random_state=rng, dtype=np.float32)
    M.data[:] = 1.0
    M.sort_indices()
    plain = rb.nmf(A, 6, seed=9, maxit=6, tol=0.0)
    zeros = rb.nmf(A, 6, seed=9, maxit=6, tol=0.0, mask="zeros")
    assert np.array_equal(plain.w, zeros.w) and np.array_equal(plain.h, zeros.h)
    masked = rb.nmf(A, 6, seed=9, maxit=6, tol=0.0, mask=M)
    assert not np.array_equal(plain.w, masked.w)
    _, _, fit = _oracle_twin(oracle, A, 6, 9, max_iter=6, tol=0.0, solver_mode=0)
    ref = fit(sort_model=True, mask=(M.indptr, M.indices))
    assert rel_err(masked.w, ref.W_T) <= RTOL and rel_err(masked.h, ref.H.T) <= RTOL and rel_err(masked.d, ref.d) <= RTOL
    empty = rb.nmf(A, 6, seed=9, maxit=6, tol=0.0, mask=sp.csc_matrix((220, 140), dtype=np.float32))
    assert np.array_equal(plain.w, empty.w)


@pytest.mark.gpu
@pytest.mark.parametrize("mask,solver", [("zeros", "cholesky"), (None, "cd")])
def test_nmf_test_fraction_runs_the_cv_entry(oracle, mask, solver):
    import rcppml_b200 as rb
    from rcppml_b200.nmf import RRandom, bridge_h_init
    m, n, k = 180, 110, 6
    A = random_csc(m, n, 0.15, 31, ragged=True)
    model = rb.nmf(A, k, seed=77, maxit=6, tol=0.0, test_fraction=0.1, mask=mask, solver=solver, cv_seed=5, cd_maxit=15,
                   sort_model=False, L1=(0.01, 0.0))
    W0 = RRandom(77).runif(m * k).reshape(k, m).T.astype(np.float32)
    H0 = bridge_h_init(77, k, n).astype(np.float32)
    ref = oracle.nmf_fit_cv(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=6, tol=0.0, L1=(0.01, 0.0),
                            solver_mode={"cd": 0, "cholesky": 1}[solver], cd_maxit=15, holdout_fraction=0.1, cv_seed=5,
                            seed=77, mask_zeros=(mask == "zeros"))
    assert model.misc["iter"] == ref.iterations and model.misc["best_iter"] == ref.best_iter + 1
    assert rel_err(model.w, ref.W_T) <= RTOL and rel_err(model.h, ref.H.T) <= RTOL and rel_err(model.d, ref.d) <= RTOL
    assert abs(model.misc["test_loss"] - ref.best_test_loss) <= 1e-5 * abs(ref.best_test_loss)
    assert abs(model.misc["train_loss"] - ref.train_loss) <= 1e-5 * abs(ref.train_loss)


@pytest.mark.gpu
def test_nnls_solves_for_h_and_for_w(oracle):
    """nnls(w =, A =) and nnls(h =, A =) (R/solve.R:300-360): c_nnls on A, resp. on the transposed problem with the
    w-side penalties; factor orientation is fixed up automatically; the old positional form still works."""
    import rcppml_b200 as rb
    m, n, k = 260, 170, 9
    A = random_csc(m, n, 0.1, 41, counts=True, ragged=True)
    rng = np.random.default_rng(3)
    w, h = rng.random((m, k)), rng.random((k, n))
    Ax = A.data.astype(np.float64)
    ref_h = oracle.project_f64(A.indptr, A.indices, Ax, m, n, w, L1=0.02, L2=0.1, upper_bound=0.5)          # (n, k)
    got_h = rb.nnls(w=w, A=A, L1=(0.3, 0.02), L2=(0.0, 0.1), upper_bound=(0.0, 0.5))
    assert got_h.shape == (k, n) and rel_err(got_h.T, ref_h) <= 1e-9
    assert np.array_equal(rb.nnls(w=w.T.copy(), A=A, L1=(0.3, 0.02), L2=(0.0, 0.1), upper_bound=(0.0, 0.5)), got_h)
    with pytest.warns(DeprecationWarning):
        old = rb.nnls(w, A)
    assert rel_err(old.T, oracle.project_f64(A.indptr, A.indices, Ax, m, n, w)) <= 1e-9
    dense = rb.nnls(w=w, A=A.toarray())
    assert np.array_equal(dense, old)
    At = A.T.tocsc()
    At.sort_indices()
    ref_w = oracle.project_f64(At.indptr, At.indices, At.data.astype(np.float64), n, m, np.ascontiguousarray(h.T),
                               L1=0.05, nonneg=False)                                                      # (m, k)
    got_w = rb.nnls(h=h, A=A, L1=(0.05, 0.4), nonneg=(False, True))
    assert got_w.shape == (m, k) and rel_err(got_w, ref_w) <= 1e-9
    h0 = rng.random((k, n)) * 0.1
    ref_ws = oracle.project_f64(A.indptr, A.indices, Ax, m, n, w, warm_start=h0.T, cd_maxit=5)
    assert rel_err(rb.nnls(w=w, A=A, warm_start=h0, cd_maxit=5).T, ref_ws) <= 1e-9


@pytest.mark.gpu
def test_nmf_multiple_initialisations_keep_the_best():
    """seed = c(5, 6, 7) / list(W1, W2) (R/nmf_thin.R:772-791, 829-925): one fit per initialisation — run i with
    config.seed = seed[1] + i - 1 — and the lowest loss wins; misc$all_inits lists them."""
    import rcppml_b200 as rb
    A = random_csc(150, 90, 0.12, 51, counts=True)
    kw = dict(maxit=4, tol=0.0, L1=0.01)
    singles = [rb.nmf(A, 5, seed=s, **kw) for s in (5, 6, 7)]
    multi = rb.nmf(A, 5, seed=[5, 6, 7], **kw)
    losses = [s.misc["loss"] for s in singles]
    best = int(np.argmin(losses))
    assert [r["loss"] for r in multi.misc["all_inits"]] == losses
    assert [r["selected"] for r in multi.misc["all_inits"]] == [i == best for i in range(3)]
    assert np.array_equal(multi.w, singles[best].w) and np.array_equal(multi.h, singles[best].h)
    assert np.array_equal(multi.misc["w_init"], singles[best].misc["w_init"]) and multi.misc["loss"] == min(losses)
    Ws = [s.misc["w_init"] for s in singles[:2]]
    from_list = rb.nmf(A, 5, seed=[Ws[0], Ws[1].T.copy()], **kw)
    assert len(from_list.misc["all_inits"]) == 2 and from_list.w.shape == (150, 5)
    assert from_list.misc["loss"] == min(r["loss"] for r in from_list.misc["all_inits"])
