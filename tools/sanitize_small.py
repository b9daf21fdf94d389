#!/usr/bin/env python
"""Tiny fits through every solve kernel (one-geometry, narrow CD, tiled; CD and Cholesky; k = 20, 64, 128), meant to run
under compute-sanitizer:   compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import rcppml_b200 as rb
    from helpers import random_csc
    eng = rb.Engine(0)
    m, n = 330, 170
    A = random_csc(m, n, 0.08, 5, ragged=True)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    ref = {}
    for k in (20, 64, 128):
        for solver in (0, 1):
            for name, env in (("wide", {"RCPPML_B200_TILED": "0", "RCPPML_B200_CD_KERNEL": "1"}),
                              ("narrow", {"RCPPML_B200_TILED": "0", "RCPPML_B200_NARROW_MIN_COLS": "0"}),
                              ("tiled", {"RCPPML_B200_TILED": "2"})):
                for kn in ("RCPPML_B200_TILED", "RCPPML_B200_CD_KERNEL", "RCPPML_B200_NARROW_MIN_COLS"):
                    os.environ.pop(kn, None)
                os.environ.update(env)
                eng.init_factors(k, 42)
                res = eng.fit(rb.make_config(k, max_iter=3, tol=0.0, solver_mode=solver, L1=(0.01, 0.01), cd_maxit=20))
                W, H, d = eng.get_factors()
                key = (k, solver)
                if key not in ref:
                    ref[key] = (W, H, d)
                same = all(np.array_equal(a, b) for a, b in zip(ref[key], (W, H, d)))
                print(f"k={k} solver={solver} {name}: status={res.status} bit-identical-to-wide={same}", flush=True)
                assert res.status == 0 and same
    eng.close()
    print("SANITIZE_RUN_OK")


if __name__ == "__main__":
    main()
