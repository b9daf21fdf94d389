#!/usr/bin/env python
"""Per-rank kernel selection for sharded runs, measured on ONE GPU.

Rank g of N runs the H half-step over n/N columns of A (gathering from the full W_T) and the W half-step over m/N rows
(gathering from the full H). Those launches are reproduced exactly by a one-GPU problem of shape m x n/N (its H
half-step) and m/N x n (its W half-step): same operands, same gather targets, no peer stores. For every N and every
kernel variant this prints the time of that half-step (CUDA-event section, mean over --steps iterations) and a digest
of the factors (all variants must agree: the kernels are bit-identical by contract).

  python tools/rank_shape_sweep.py [--ns 1,2,4,8] [--k 64] [--solver 1] [--out gpurun_out/rank_shapes.jsonl]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KNOBS = ("RCPPML_B200_TILED", "RCPPML_B200_TILED_SL", "RCPPML_B200_TILED_CTA", "RCPPML_B200_NV", "RCPPML_B200_NV_SHORT", "RCPPML_B200_CD_GEOM",
         "RCPPML_B200_CD_KERNEL", "RCPPML_B200_TILED_MIN_BATCHES", "RCPPML_B200_NARROW_MIN_COLS", "RCPPML_B200_PANEL_MB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--density", type=float, default=1e-3)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--solver", type=int, default=1)
    ap.add_argument("--ns", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default="")
    ap.add_argument("--variants", default="", help="comma-separated subset of the variant names")
    ap.add_argument("--half-steps", default="H,W")
    args = ap.parse_args()

    import torch
    import rcppml_b200 as rb

    torch.cuda.set_device(0)
    variants = [("untiled", {"RCPPML_B200_TILED": "0"}),
                ("tiled16", {"RCPPML_B200_TILED": "2", "RCPPML_B200_TILED_SL": "2"}),
                ("tiled8", {"RCPPML_B200_TILED": "2", "RCPPML_B200_TILED_SL": "4"}),
                ("tiled4", {"RCPPML_B200_TILED": "2", "RCPPML_B200_TILED_SL": "8"}),
                ("tiled8_cta768", {"RCPPML_B200_TILED": "2", "RCPPML_B200_TILED_SL": "4", "RCPPML_B200_TILED_CTA": "1"}),
                ("tiled8_cta768_hybrid", {"RCPPML_B200_TILED": "2", "RCPPML_B200_TILED_SL": "4", "RCPPML_B200_TILED_CTA": "2"}),
                ("untiled_nv1", {"RCPPML_B200_TILED": "0", "RCPPML_B200_NV": "1"}),
                ("untiled_nv2", {"RCPPML_B200_TILED": "0", "RCPPML_B200_NV": "2"}),
                ("default", {})]
    if args.variants:
        keep = set(args.variants.split(","))
        variants = [v for v in variants if v[0] in keep]
    lines = []
    for N in [int(x) for x in args.ns.split(",")]:
        for which, (m_s, n_s) in (("H", (args.m, args.n // N)), ("W", (args.m // N, args.n))):
            if which not in args.half_steps.split(","):
                continue
            eng = rb.Engine(0)
            # columns [0, n_s) / all columns of the generator with m_s rows: same column statistics as the rank's operand
            eng.set_matrix_synthetic(m_s, n_s, 0, args.density, 20260101)
            digests = set()
            for name, env in variants:
                for kname in KNOBS:
                    os.environ.pop(kname, None)
                os.environ.update(env)
                eng.init_factors(args.k, 42, 0)
                cfg = rb.make_config(args.k, max_iter=args.steps + args.warmup, tol=0.0, solver_mode=args.solver, cd_maxit=100)
                eng.set_profiling(False)
                eng.begin_fit(cfg)
                eng.iterate(args.warmup)
                eng.set_profiling(True)
                torch.cuda.synchronize()
                eng.iterate(args.steps)
                torch.cuda.synchronize()
                res = eng.result()
                prof_ms, _ = eng.profile()
                cs = eng.factor_checksum()
                digests.add(cs)
                sec = "fused_rhs_nnls_H" if which == "H" else "fused_rhs_nnls_W"
                line = {"N": N, "half_step": which, "shape": [m_s, n_s], "nnz": eng.nnz, "variant": name, "env": env,
                        "k": args.k, "solver_mode": args.solver, "half_step_ms": prof_ms[sec] / args.steps,
                        "iteration_ms": res.loop_ms / args.steps, "status": res.status,
                        "checksum": "%016x" % cs[0]}
                lines.append(line)
                print(json.dumps(line), flush=True)
            eng.close()
            print(json.dumps({"N": N, "half_step": which, "all_variants_bit_identical": len(digests) == 1}), flush=True)
    for kname in KNOBS:
        os.environ.pop(kname, None)
    if args.out:
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
