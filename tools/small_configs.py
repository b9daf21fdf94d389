#!/usr/bin/env python
"""Per-iteration time of the small BASELINE.json configurations (C1 aml, C2 movielens, C3 pbmc3k) on one GPU, with the
steady-state iteration replayed as a CUDA graph (default) and with plain launches (RCPPML_B200_GRAPH=0).

  python tools/small_configs.py [--iters 200] [--out gpurun_out/small_configs.jsonl]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--out", default="")
    ap.add_argument("--kernel-variants", action="store_true",
                    help="also time the other kernel selections (RCPPML_B200_TILED / _CD_KERNEL knobs)")
    args = ap.parse_args()
    import torch
    import rcppml_b200 as rb
    from helpers import load_aml_as_csc, load_movielens, load_pbmc3k

    torch.cuda.set_device(0)
    eng = rb.Engine(0)
    cases = [("C1 aml 824x135 (as CSC) k=6", load_aml_as_csc(), 6, {}),
             ("C2 movielens 3867x610 k=20 L1=0.01", load_movielens()[0], 20, dict(L1=(0.01, 0.01)))]
    pb = load_pbmc3k()
    if pb is not None:
        cases.append(("C3 pbmc3k 13714x2700 k=32", pb, 32, {}))
    lines = []
    for name, A, k, kw in cases:
        m, n = A.shape
        eng.set_matrix(m, n, A.indptr, A.indices, A.data)
        for solver in (0, 1):
            row = {"config": name, "nnz": int(A.nnz), "k": k, "solver_mode": solver, "iters": args.iters}
            for graph in ("1", "0"):
                os.environ["RCPPML_B200_GRAPH"] = graph
                best = None
                for rep in range(3):
                    eng.init_factors(k, 42)
                    cfg = rb.make_config(k, max_iter=args.iters, tol=0.0, solver_mode=solver, **kw)
                    t0 = time.perf_counter()
                    res = eng.fit(cfg)
                    wall = time.perf_counter() - t0
                    assert res.status == 0 and res.iterations == args.iters
                    cur = (res.loop_ms / args.iters, 1e3 * wall / args.iters)
                    best = cur if best is None or cur[0] < best[0] else best
                row["graph" if graph == "1" else "plain"] = {"device_ms_per_iter": best[0], "wall_ms_per_iter": best[1]}
            row["nnz_per_sec_graph"] = A.nnz / (row["graph"]["device_ms_per_iter"] / 1e3)
            if args.kernel_variants:
                os.environ["RCPPML_B200_GRAPH"] = "1"
                variants = {"tiled_always": {"RCPPML_B200_TILED": "2"}, "untiled": {"RCPPML_B200_TILED": "0"}}
                if solver == 0:
                    variants["untiled_wide_cd"] = {"RCPPML_B200_TILED": "0", "RCPPML_B200_CD_KERNEL": "1"}
                for vname, env in variants.items():
                    os.environ.update(env)
                    best = None
                    for rep in range(2):
                        eng.init_factors(k, 42)
                        res = eng.fit(rb.make_config(k, max_iter=args.iters, tol=0.0, solver_mode=solver, **kw))
                        best = res.loop_ms / args.iters if best is None else min(best, res.loop_ms / args.iters)
                    row[vname] = best
                    for kn in env:
                        os.environ.pop(kn, None)
            lines.append(row)
            print(json.dumps(row), flush=True)
    os.environ.pop("RCPPML_B200_GRAPH", None)
    eng.close()
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
