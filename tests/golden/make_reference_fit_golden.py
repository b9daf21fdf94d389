"""Golden vectors FROM THE REFERENCE ITSELF: runs the reference's own nmf_fit / nmf_fit_cv (oracle/_ref/libref_fit.so —
nmf/fit_cpu.hpp and nmf/fit_cv.hpp compiled unmodified against the Eigen stand-in, `make -C oracle ref_hotpath`) on
small seeded inputs and stores inputs and outputs in tests/golden/reference_fits.npz, so that the oracle (CPU suite)
and the CUDA engine (GPU suite) can be held against the reference's answers wherever the library itself is not
available.

  python tests/golden/make_reference_fit_golden.py        (needs /root/reference; run in the build container)
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import random_csc  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_reference_fit import CP, CR, R, P, _p, _ref_fit  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_fit.so"))
out = {}
FITS = [("cd_k8", 120, 90, 0.15, 8, dict(solver_mode=0)),
        ("chol_k8_L1L2", 120, 90, 0.15, 8, dict(solver_mode=1, L1=(0.01, 0.02), L2=(0.02, 0.01))),
        ("cd_k20_L1", 150, 100, 0.12, 20, dict(solver_mode=0, L1=(0.01, 0.01))),
        ("chol_k32", 160, 110, 0.12, 32, dict(solver_mode=1)),
        ("chol_k6_ub_l2norm", 100, 80, 0.2, 6, dict(solver_mode=1, upper_bound=(0.3, 0.2), norm_type=1)),
        ("cd_k7_seminmf", 100, 80, 0.2, 7, dict(solver_mode=0, nonneg=(True, False)))]
names = []
for name, m, n, dens, k, kw in FITS:
    A = random_csc(m, n, dens, 23, counts=("L1" in name), ragged=True)
    W0, H0 = O.initialize_factors(k, m, n, 42)
    W, H, d, hist, res = _ref_fit(lib, A, k, W0, H0, max_iter=6, **kw)
    for key, val in dict(indptr=A.indptr.astype(np.int32), indices=A.indices.astype(np.int32), data=A.data.astype(np.float32),
                         shape=np.array([m, n, k, 6]), W0=W0, H0=H0, W=W, H=H, d=d, loss=hist,
                         kw=np.array(repr(kw))).items():
        out[f"fit/{name}/{key}"] = val
    names.append(name)
out["fit_names"] = np.array(names)

cv_names = []
for name, k, solver, mz in (("cv_chol_mz", 7, 1, True), ("cv_cd_mz", 7, 0, True), ("cv_chol_all", 5, 1, False)):
    m, n, iters = 110, 80, 5
    A = random_csc(m, n, 0.2, 29, ragged=True)
    W0, H0 = O.initialize_factors(k, m, n, 42)
    q = P(k=k, max_iter=iters, tol=0.0, L1_W=0.01, L1_H=0.0, L2_W=0.0, L2_H=0.01, ub_W=0, ub_H=0, nonneg_W=1, nonneg_H=1,
          cd_maxit=15, cd_tol=1e-8, norm_type=0, solver_mode=solver, patience=5, threads=1, sort_model=0, seed=42)
    cq = CP(holdout_fraction=0.1, cv_seed=7, mask_zeros=int(mz), cv_patience=5)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float32)
    W_in, H_in = np.ascontiguousarray(W0.T), np.ascontiguousarray(H0)
    W_out, H_out, d = np.zeros((k, m), np.float32), np.zeros((n, k), np.float32), np.zeros(k, np.float32)
    tr, te = np.zeros(iters, np.float32), np.zeros(iters, np.float32)
    res, cres, err = R(), CR(), C.create_string_buffer(300)
    rc = lib.reffit_nmf_cv_sparse_f32(_p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_float), m, n, C.byref(q), C.byref(cq),
                                      _p(W_in, C.c_float), _p(H_in, C.c_float), _p(W_out, C.c_float), _p(H_out, C.c_float),
                                      _p(d, C.c_float), _p(tr, C.c_float), _p(te, C.c_float), C.byref(res), C.byref(cres), err, 300)
    assert rc == 0, err.value
    for key, val in dict(indptr=Ap, indices=Ai, data=Ax, shape=np.array([m, n, k, iters, solver, int(mz)]), W0=W0, H0=H0,
                         W=W_out.T.copy(), H=H_out, d=d, train=tr[:res.n_loss], test=te[:cres.n_test_hist],
                         best_iter=np.array(cres.best_iter)).items():
        out[f"cv/{name}/{key}"] = val
    cv_names.append(name)
out["cv_names"] = np.array(cv_names)
np.savez_compressed(os.path.join(HERE, "reference_fits.npz"), **out)
print("wrote reference_fits.npz:", names, cv_names, os.path.getsize(os.path.join(HERE, "reference_fits.npz")), "bytes")
