#!/usr/bin/env python
"""Timings of the SURVEY.md §8f rows at the C4 size on one GPU (host-call wall clock, data starting in host memory
unless noted): speckled-mask cross-validation fit (device loop time), predict() (one fp64 projection of all columns),
evaluate() over the non-zeros and over all m*n entries.

  python tools/next_rows_bench.py [--m 1000000 --n 100000 --k 64] [--out gpurun_out/next_rows.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--density", type=float, default=1e-3)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--cv-iters", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import rcppml_b200 as rb
    from rcppml_b200 import project

    torch.cuda.set_device(0)
    eng = rb.Engine(0)
    eng.set_matrix_synthetic_sharded(args.m, args.n, args.density, 20260101)
    nnz = eng.nnz_global
    out = {"workload": f"synthetic {args.m}x{args.n} density {args.density:g} (nnz {nnz}), k={args.k}", "rows": {}}

    # ---- §8f-1: cross-validation fit (speckled mask in-kernel, per-column Gram downdates)
    for solver, name in ((1, "cholesky"), (0, "cd")):
        eng.init_factors(args.k, 42, 0)
        cfg = rb.make_config(args.k, max_iter=args.cv_iters, tol=0.0, solver_mode=solver, cd_maxit=100)
        t0 = time.perf_counter()
        res, cv = eng.fit_cv(cfg, holdout_fraction=0.1, cv_seed=7, mask_zeros=True)
        wall = time.perf_counter() - t0
        out["rows"][f"cv_fit_{name}"] = {"iterations": res.iterations, "device_ms_per_iter": res.loop_ms / max(1, res.iterations),
                                         "wall_s": wall, "n_test": int(cv["n_test"]), "test_loss": float(cv["test_loss"]),
                                         "train_loss": float(cv["train_loss"])}
        print(json.dumps({f"cv_fit_{name}": out["rows"][f"cv_fit_{name}"]}), flush=True)

    # ---- §8f-2 / -3: predict and evaluate through the fp64 entry points (host CSC in, host result out)
    p, i, x = eng.get_matrix()
    W, H, d = eng.get_factors()
    eng.close()
    A = (p, i, x.astype(np.float64), (args.m, args.n))
    w64 = W.astype(np.float64)
    for rep in range(2):
        t0 = time.perf_counter()
        h = project.predict(w64, A)
        t_pred = time.perf_counter() - t0
    out["rows"]["predict"] = {"wall_s": t_pred, "nnz_per_s": nnz / t_pred, "h_shape": list(h.shape)}
    print(json.dumps({"predict": out["rows"]["predict"]}), flush=True)
    for mz in (True, False):
        for rep in range(2):
            t0 = time.perf_counter()
            mse = project.evaluate(A, w64, d.astype(np.float64), H.astype(np.float64).T, mask_zeros=mz)
            t_ev = time.perf_counter() - t0
        out["rows"][f"evaluate_mask_zeros_{mz}"] = {"wall_s": t_ev, "mse": mse}
        print(json.dumps({f"evaluate_mask_zeros_{mz}": out["rows"][f"evaluate_mask_zeros_{mz}"]}), flush=True)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
