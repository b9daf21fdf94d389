"""Generates tests/golden/oracle_small_fit.npz from the CPU oracle (NOT from the reference: the reference
needs Eigen/R, absent from this image — SURVEY.md §8c). The fixture freezes the restatement so that a
later edit of oracle/nmf_oracle.cpp cannot silently change what the GPU is compared against.

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import random_csc  # noqa: E402
from oracle import oracle as O  # noqa: E402

m, n, k, iters = 90, 70, 7, 6
A = random_csc(m, n, 0.15, 2024, counts=True, ragged=True)
W0, H0 = O.initialize_factors(k, m, n, 42)
out = dict(indptr=A.indptr, indices=A.indices, data=A.data, m=m, n=n, k=k, iters=iters, W0=W0, H0=H0)
for solver in (0, 1):
    r = O.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver,
                  L1=(0.01, 0.02), L2=(0.0, 0.01), threads=1)
    out[f"W_{solver}"], out[f"H_{solver}"], out[f"d_{solver}"], out[f"loss_{solver}"] = r.W_T, r.H, r.d, r.loss_history
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small_fit.npz"), **out)
print("wrote oracle_small_fit.npz")

# pbmc3k[0:500, 0:200] — the block of the reference's GPU accuracy test (tests/testthat/test_gpu_accuracy.R:26-34),
# cut from the matrix the reference's own .spz decoder produced (oracle/_ref/pbmc3k.bin, `make -C oracle ref`).
from helpers import load_pbmc3k  # noqa: E402
A = load_pbmc3k()
if A is not None:
    B = A[:500, :200].tocsc()
    B.sort_indices()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pbmc3k_500x200.npz"),
                        indptr=B.indptr.astype(np.int32), indices=B.indices.astype(np.int32), data=B.data.astype(np.float32))
    print("wrote pbmc3k_500x200.npz", B.shape, B.nnz)
