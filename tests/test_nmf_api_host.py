"""Host logic of rcppml_b200.nmf() / nnls() / predict() / evaluate() on a box WITHOUT a GPU.

The bodies of the `-m gpu` tests in test_nmf_api.py are run once more with the native calls (the ctypes bridge twins and
the projection entry points) replaced by stand-ins built on the CPU oracle. What this covers is everything ABOVE the C
ABI: argument mapping and validation, (W, H) pair order, orientation of w / h, R's RNG and the bridge's H stream, CV
and mask routing, sort_model, result packing. It says nothing about the CUDA path — the same bodies check that on the
GPU box. The stand-ins live in this test file only; the product package never imports the oracle."""
import sys

import numpy as np
import pytest

import test_nmf_api as T


@pytest.fixture
def oracle_backed_native_calls(monkeypatch, oracle):
    import rcppml_b200  # noqa: F401
    import rcppml_b200.bridge as B
    import rcppml_b200.project as P
    nm = sys.modules["rcppml_b200.nmf"]
    calls = []

    def fake_sparse(indptr, indices, data, m, n, k, W_T0, H0, *, mask=None, seed=0, verbose=False, **kw):
        calls.append(("masked" if mask is not None else "standard", dict(kw, seed=seed)))
        r = oracle.nmf_fit(indptr, indices, data, m, n, k, np.asarray(W_T0, np.float32),
                           np.asarray(H0, np.float64).astype(np.float32), mask=mask, **kw)
        return B.BridgeResult(r.W_T, r.H, r.d, r.iterations, r.converged, r.train_loss, r.final_tol, 0)

    def fake_cv(indptr, indices, data, m, n, k, W_T0, H0, *, verbose=False, loss_type=0, projective=False,
                symmetric=False, **kw):
        calls.append(("cv", dict(kw)))
        r = oracle.nmf_fit_cv(indptr, indices, data, m, n, k, np.asarray(W_T0, np.float32),
                              np.asarray(H0).astype(np.float32), **kw)
        return B.BridgeCvResult(r.W_T, r.H, r.d, r.iterations, r.converged, r.train_loss, r.test_loss, r.best_test_loss,
                                r.best_iter, 0)

    def fake_nnls(w, A, *, L1=0.0, L2=0.0, upper_bound=0.0, nonneg=True, cd_maxit=100, cd_tol=1e-8, warm_start=None):
        ip, ii, dd, (m, n) = P._csc(A)
        ws = None if warm_start is None else np.ascontiguousarray(np.asarray(warm_start, np.float64).T)
        return oracle.project_f64(ip, ii, dd, m, n, np.ascontiguousarray(w, np.float64), L1=L1, L2=L2,
                                  upper_bound=upper_bound, nonneg=nonneg, cd_maxit=cd_maxit, cd_tol=cd_tol,
                                  warm_start=ws).T.copy()

    def fake_evaluate(A, w, d, h, *, mask_zeros=False):
        ip, ii, dd, (m, n) = P._csc(A)
        return oracle.evaluate_mse_f64(ip, ii, dd, m, n, np.ascontiguousarray(w, np.float64), np.asarray(d, np.float64),
                                       np.ascontiguousarray(np.asarray(h, np.float64).T), mask_zeros)

    monkeypatch.setattr(nm, "bridge_nmf_sparse", fake_sparse)
    monkeypatch.setattr(nm, "bridge_nmf_cv_sparse", fake_cv)
    monkeypatch.setattr(nm, "gpu_detect", lambda: {"status": 0, "num_gpus": 1})
    monkeypatch.setattr(P, "nnls", fake_nnls)
    monkeypatch.setattr(P, "predict", lambda w, A, *, L1=0.0, L2=0.0, upper_bound=0.0: fake_nnls(
        w, A, L1=L1, L2=L2, upper_bound=upper_bound))
    monkeypatch.setattr(P, "evaluate", fake_evaluate)
    return calls


def _params(fn):
    return [m for m in fn.pytestmark if m.name == "parametrize"][0].args[1]


def test_nmf_host_plumbing(oracle_backed_native_calls, oracle):
    calls = oracle_backed_native_calls
    for k, kw, okw in _params(T.test_nmf_matches_the_oracle):
        T.test_nmf_matches_the_oracle(oracle, k, kw, okw)
    # what reached the bridge for nmf(A, 40, L2=(0.01, 0.0)): solver auto -> Cholesky above k = 32, pairs as (W, H)
    sent = [c for c in calls if c[0] == "standard" and c[1].get("L2") == (0.01, 0.0)]
    assert sent and all(c[1]["solver_mode"] == 1 and c[1]["seed"] == 123 and c[1]["cd_maxit"] == 100 for c in sent)
    T.test_nmf_seed_forms_and_reproducibility()
    T.test_nmf_invariants_of_the_reference_tests()
    calls.clear()
    T.test_nmf_multiple_initialisations_keep_the_best()
    assert [c[1]["seed"] for c in calls[3:6]] == [5, 6, 7]             # config.seed = seed[1] + i - 1


def test_nmf_mask_and_cv_routing(oracle_backed_native_calls, oracle):
    calls = oracle_backed_native_calls
    T.test_nmf_explicit_mask_and_mask_zeros(oracle)
    kinds = [c[0] for c in calls]
    assert kinds.count("masked") == 1 and "cv" not in kinds            # "zeros" / an empty mask stay on the standard entry
    calls.clear()
    for mask, solver in _params(T.test_nmf_test_fraction_runs_the_cv_entry):
        T.test_nmf_test_fraction_runs_the_cv_entry(oracle, mask, solver)
    assert [c[0] for c in calls] == ["cv", "cv"]
    assert calls[0][1]["mask_zeros"] is True and calls[1][1]["mask_zeros"] is False
    assert all(abs(c[1]["holdout_fraction"] - float(np.float32(0.1))) == 0 and c[1]["cv_seed"] == 5 for c in calls)


def test_nnls_host_plumbing(oracle_backed_native_calls, oracle):
    T.test_nnls_solves_for_h_and_for_w(oracle)
