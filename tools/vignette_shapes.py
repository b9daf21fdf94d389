#!/usr/bin/env python
"""The shapes of the reference's own GPU vignette (vignettes/gpu-acceleration.Rmd:82-96: end-to-end `system.time` of
nmf(), tol = 0, maxit = 20, on one H100 NVL vs a 56-thread Xeon — BASELINE.md §1), run end to end through
rcppml_gpu_nmf_unified_float on this GPU from host memory: random sparse matrices of the published size and density
(the vignette's data are simulated too), k and iteration count as published. For scale only — different hardware,
different random matrix; the published seconds are quoted beside ours.

  python tools/vignette_shapes.py [--out gpurun_out/vignette_shapes.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PUBLISHED = [  # (label, m, n, nnz, k, iters, published CPU s, published GPU s)
    ("NMF k=20, 5000x40000, nnz 33M", 5000, 40000, 33_000_000, 20, 20, 38.45, 2.78),
    ("NMF k=64, 5000x10000, nnz 8.3M", 5000, 10000, 8_300_000, 64, 20, 29.23, 0.88),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import rcppml_b200 as rb
    rows = []
    for label, m, n, nnz, k, iters, cpu_s, gpu_s in PUBLISHED:
        rng = np.random.default_rng(1)
        A = sp.random(m, n, density=nnz / (m * n), format="csc", random_state=rng, dtype=np.float32)
        A.data = (A.data + 0.5).astype(np.float32)
        A.sort_indices()
        W0, H0 = rng.random((m, k)), rng.random((n, k))
        for solver, name in ((1, "cholesky (R's GPU default at k > 32)"), (0, "cd (R's default at k <= 32)")):
            best = None
            for rep in range(3):                       # first call: allocations; report the best of the next two
                t0 = time.perf_counter()
                out = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0,
                                           solver_mode=solver, cd_maxit=100)
                dt = time.perf_counter() - t0
                assert out.status == 0 and out.iterations == iters
                if rep > 0:
                    best = dt if best is None else min(best, dt)
            rows.append({"shape": label, "nnz": int(A.nnz), "k": k, "iterations": iters, "solver": name,
                         "b200_end_to_end_s": best, "published_h100_gpu_s": gpu_s, "published_xeon56_cpu_s": cpu_s,
                         "note": "host numpy arrays (pageable), doubles on the wire, incl. the host-side packing of bridge.py"})
            print(json.dumps(rows[-1]), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
