"""Golden vectors produced by THE REFERENCE ITSELF (tests/golden/reference_fits.npz, written by
tests/golden/make_reference_fit_golden.py from the reference's own nmf_fit / nmf_fit_cv compiled against the Eigen
stand-in). They need neither /root/reference nor oracle/_ref at test time:
  * CPU: the oracle reproduces the reference's W, d, H bit for bit (loss to 1e-5);
  * GPU: the CUDA engine matches them to 1e-5 with identical zero patterns."""
import ast
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import RTOL, rel_err, zero_pattern_equal

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_fits.npz"))
FITS = [str(x) for x in G["fit_names"]]
CVS = [str(x) for x in G["cv_names"]]


def _fit_case(name):
    g = {k.split("/")[-1]: G[k] for k in G.files if k.startswith(f"fit/{name}/")}
    m, n, k, iters = (int(v) for v in g["shape"])
    return g, m, n, k, iters, ast.literal_eval(str(g["kw"]))


@pytest.mark.parametrize("name", FITS)
def test_oracle_matches_reference_golden(oracle, name):
    g, m, n, k, iters, kw = _fit_case(name)
    r = oracle.nmf_fit(g["indptr"], g["indices"], g["data"], m, n, k, g["W0"], g["H0"], max_iter=iters, tol=0.0, threads=1, **kw)
    assert np.array_equal(r.W_T, g["W"]) and np.array_equal(r.H, g["H"]) and np.array_equal(r.d, g["d"])
    assert np.allclose(r.loss_history, g["loss"], rtol=1e-5, atol=0)


@pytest.mark.parametrize("name", CVS)
def test_oracle_cv_matches_reference_golden(oracle, name):
    g = {k.split("/")[-1]: G[k] for k in G.files if k.startswith(f"cv/{name}/")}
    m, n, k, iters, solver, mz = (int(v) for v in g["shape"])
    r = oracle.nmf_fit_cv(g["indptr"], g["indices"], g["data"], m, n, k, g["W0"], g["H0"], max_iter=iters, tol=0.0,
                          solver_mode=solver, L1=(0.01, 0.0), L2=(0.0, 0.01), cd_maxit=15, holdout_fraction=0.1, cv_seed=7,
                          seed=42, mask_zeros=bool(mz), threads=1)
    assert np.array_equal(r.W_T, g["W"]) and np.array_equal(r.H, g["H"]) and np.array_equal(r.d, g["d"])
    assert r.best_iter == int(g["best_iter"])
    assert np.allclose(r.test_history, g["test"], rtol=1e-5, atol=0) and np.allclose(r.train_history, g["train"], rtol=1e-5, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FITS)
def test_gpu_matches_reference_golden(name):
    import rcppml_b200 as rb
    g, m, n, k, iters, kw = _fit_case(name)
    eng = rb.Engine(0)
    try:
        eng.set_matrix(m, n, g["indptr"], g["indices"], g["data"])
        eng.set_factors(g["W0"], g["H0"])
        res = eng.fit(rb.make_config(k, max_iter=iters, tol=0.0, **kw))
        W, H, d = eng.get_factors()
        loss = eng.loss_history(iters)
    finally:
        eng.close()
    assert res.status == 0 and res.iterations == iters
    errs = dict(W=rel_err(W, g["W"]), H=rel_err(H, g["H"]), d=rel_err(d, g["d"]), loss=rel_err(loss, g["loss"]))
    assert max(errs.values()) <= RTOL, (name, errs)
    assert zero_pattern_equal(W, g["W"]) and zero_pattern_equal(H, g["H"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CVS)
def test_gpu_cv_matches_reference_golden(name):
    import rcppml_b200 as rb
    g = {k.split("/")[-1]: G[k] for k in G.files if k.startswith(f"cv/{name}/")}
    m, n, k, iters, solver, mz = (int(v) for v in g["shape"])
    eng = rb.Engine(0)
    try:
        eng.set_matrix(m, n, g["indptr"], g["indices"], g["data"])
        eng.set_factors(g["W0"], g["H0"])
        cfg = rb.make_config(k, max_iter=iters, tol=0.0, solver_mode=solver, L1=(0.01, 0.0), L2=(0.0, 0.01), cd_maxit=15)
        res, cv = eng.fit_cv(cfg, holdout_fraction=0.1, cv_seed=7, seed=42, mask_zeros=bool(mz))
        W, H, d = eng.get_factors()
    finally:
        eng.close()
    assert res.status == 0 and res.iterations == iters and cv["best_iter"] == int(g["best_iter"])
    assert rel_err(W, g["W"]) <= RTOL and rel_err(H, g["H"]) <= RTOL and rel_err(d, g["d"]) <= RTOL
    assert np.allclose(cv["test_history"], g["test"], rtol=1e-5, atol=0)
