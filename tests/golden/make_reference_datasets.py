"""Extracts the reference's own datasets for BASELINE.json configs[0] and configs[1] into committed fixtures:

  data/movielens.rda (Matrix::dgCMatrix 3867 x 610, 75 238 ratings)  -> tests/golden/movielens.npz
  data/aml.rda       (dense 824 x 135 methylation matrix)            -> tests/golden/aml.npz

The .rda files are R `save()` images; tests/golden/rdx3.py reads them without R. Also freezes what the CPU oracle
makes of them (loss history, d, CD sweep totals) so that the oracle itself is pinned on real data.

  python tests/golden/make_reference_datasets.py        (needs /root/reference; run in the build container)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import rdx3  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/data"

# ---- movielens: configs[1] "movielens sparse dgCMatrix k=20 L1=0.01"
obj = rdx3.read_rda(os.path.join(REF, "movielens.rda"))["movielens"]
p, i, x, (m, n) = rdx3.as_csc(obj)
assert (m, n) == (3867, 610) and p[-1] == len(i) == len(x)
assert all(np.all(np.diff(i[p[j]:p[j + 1]]) > 0) for j in range(n)), "row indices must be ascending (dgCMatrix)"
x32 = x.astype(np.float32)
k, iters = 20, 6
W0, H0 = O.initialize_factors(k, m, n, 42)
out = dict(indptr=p, indices=i, data=x32, m=m, n=n)
for solver in (0, 1):
    r = O.nmf_fit(p, i, x32, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver, L1=(0.01, 0.01), threads=1)
    out[f"loss_{solver}"], out[f"d_{solver}"], out[f"sweeps_{solver}"] = r.loss_history, r.d, r.cd_sweeps
np.savez_compressed(os.path.join(HERE, "movielens.npz"), **out)
print("wrote movielens.npz", (m, n), len(x), "ratings", sorted(set(np.round(x, 2)))[:12])

# ---- aml: configs[0] "aml dense 824x135 k=6 loss=mse" (the reference quick-start; dense, CPU only)
obj = rdx3.read_rda(os.path.join(REF, "aml.rda"))["aml"]
dim = np.asarray(obj.attrs["dim"])
A = np.asarray(obj.value, np.float64).reshape((int(dim[1]), int(dim[0]))).T          # column-major -> (824, 135)
assert A.shape == (824, 135)
A32 = np.ascontiguousarray(A.astype(np.float32))
np.savez_compressed(os.path.join(HERE, "aml.npz"), A=A32)
print("wrote aml.npz", A32.shape, "min/max", float(A32.min()), float(A32.max()), "zeros", int((A32 == 0).sum()))
