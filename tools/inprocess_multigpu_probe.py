"""C4 through rcppml_gpu_nmf_unified_float with RCPPML_NUM_GPUS = 1, 2, ... (single process, one host thread per
device inside the library): wall time of the whole call from pinned host buffers, the engine's phase times, and a
bit-for-bit comparison of the factors with the one-GPU call.

    python tools/inprocess_multigpu_probe.py [--gpus 1,2,4,8] [--iters 20] [--out gpurun_out/inprocess_multigpu.json]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcppml_b200 as rb  # noqa: E402
from rcppml_b200 import _lib, bridge, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", default="1,2")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--m", type=int, default=1_000_000)
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--k", type=int, default=64)
ap.add_argument("--solver", type=int, default=1)
ap.add_argument("--out", default="")
a = ap.parse_args()

import torch  # noqa: E402

m, n, k = a.m, a.n, a.k
eng = rb.Engine(0)
eng.set_matrix_synthetic_sharded(m, n, 1e-3, synth.SEED_A)
eng.init_factors(k, 42, 0)
p, i, x = eng.get_matrix()
W0, H0, _ = eng.get_factors()
eng.close()
pin = lambda arr: torch.from_numpy(arr).pin_memory().numpy()
p, i, x64 = pin(p), pin(i), pin(x.astype(np.float64))
nnz = int(p[n])
rows, ref = [], None
for G in [int(g) for g in a.gpus.split(",")]:
    os.environ["RCPPML_NUM_GPUS"] = str(G)
    best = None
    for rep in range(3):                                   # first call warms the context(s) up
        W, H = pin(W0.astype(np.float64)), pin(H0.astype(np.float64))
        call = bridge.PackedCall(p, i, x64, m, n, k, W, H, max_iter=a.iters, tol=0.0, solver_mode=a.solver)
        t = time.perf_counter()
        call()
        dt = time.perf_counter() - t
        assert call.status == 0, (G, call.status)
        ph = (C.c_double * 5)()
        _lib.load().rcppml_b200_last_call_phases(ph)
        if rep and (best is None or dt < best["seconds"]):
            best = dict(gpus=G, seconds=dt, nnz_per_sec=nnz * a.iters / dt, iterations=call.iterations,
                        phases_ms=dict(zip(("matrix_h2d", "transpose", "factors_h2d", "als_loop", "factors_d2h"), list(ph))))
    if ref is None:
        ref = (W.copy(), H.copy(), call.d.copy(), call.train_loss)
        best["bit_identical_to_first"] = True
    else:
        best["bit_identical_to_first"] = bool(np.array_equal(ref[0], W) and np.array_equal(ref[1], H)
                                              and np.array_equal(ref[2], call.d) and ref[3] == call.train_loss)
    print(json.dumps(best), flush=True)
    rows.append(best)
if a.out:
    with open(a.out, "w") as f:
        json.dump(rows, f, indent=1)
