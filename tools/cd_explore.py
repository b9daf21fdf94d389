#!/usr/bin/env python
"""Coordinate-descent (solver_mode 0) geometry sweep on the C4 workload: times a few ALS iterations per
lane-group geometry of cd_half_step_kernel and of the original half_step_kernel<CD>, checks that every
variant produces bit-identical factors and sweep counts, prints one JSON line per variant.

  python tools/cd_explore.py [--m 1000000 --n 100000 --k 64 --steps 3 --warmup 1] [--out gpurun_out/cd.jsonl]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--density", type=float, default=1e-3)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--solver", type=int, default=0)
    ap.add_argument("--variants", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    import rcppml_b200 as rb

    torch.cuda.set_device(0)
    eng = rb.Engine(0)
    eng.set_matrix_synthetic_sharded(args.m, args.n, args.density, 20260101)
    kp = 16 if args.k <= 16 else 32 if args.k <= 32 else 64 if args.k <= 64 else 128
    geoms = {16: (301, 102, 4), 32: (701, 302, 104), 64: (702, 304, 108), 128: (704, 308, 116)}[kp]
    if args.solver == 0:
        variants = [("default", {}), ("untiled", {"RCPPML_B200_TILED": "0"})]
        variants += [("v2_geom%d" % g, {"RCPPML_B200_CD_GEOM": str(g), "RCPPML_B200_TILED": "0"}) for g in geoms]
        variants += [("v1_nv%d" % nv, {"RCPPML_B200_CD_KERNEL": "1", "RCPPML_B200_NV": str(nv)})
                     for nv in ((1, 2, 4) if kp >= 64 else (1,))]
    else:
        variants = [("chol_default", {}), ("chol_untiled", {"RCPPML_B200_TILED": "0"}), ("chol_tiled_both", {"RCPPML_B200_TILED": "2"})]
        if kp >= 64:
            variants += [("chol_untiled_nvshort%d" % nv, {"RCPPML_B200_NV_SHORT": str(nv), "RCPPML_B200_TILED": "0"}) for nv in (1, 4)]
            variants += [("chol_tiled_gather16x1", {"RCPPML_B200_NV_SHORT": "1"})]
    if args.variants:
        keep = set(args.variants.split(","))
        variants = [v for v in variants if v[0] in keep]
    knobs = ("RCPPML_B200_CD_GEOM", "RCPPML_B200_CD_KERNEL", "RCPPML_B200_NV", "RCPPML_B200_NV_SHORT", "RCPPML_B200_TILED")
    lines = []
    for name, env in variants:
        for kname in knobs:
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
        W, H, d = eng.get_factors()
        digest = hashlib.sha256(W.tobytes() + H.tobytes() + d.tobytes()).hexdigest()[:16]
        line = {"variant": name, "env": env, "k": args.k, "solver_mode": args.solver,
                "ms_per_iter": res.loop_ms / args.steps, "status": res.status,
                "cd_sweeps_total": eng.cd_sweeps(), "digest": digest,
                "sections_ms_per_iter": {kk: v / args.steps for kk, v in prof_ms.items()}}
        lines.append(line)
        print(json.dumps(line), flush=True)
    eng.close()
    ok = len({(l["digest"], l["cd_sweeps_total"]) for l in lines}) == 1
    print(json.dumps({"all_variants_bit_identical": ok}), flush=True)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")
            f.write(json.dumps({"all_variants_bit_identical": ok}) + "\n")


if __name__ == "__main__":
    main()
