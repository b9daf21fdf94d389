#!/usr/bin/env python
"""File -> device -> fit on one GPU: the on-disk ingest end to end. Writes a synthetic matrix as a .spz v2
file with the REFERENCE's writer (oracle/_ref/spz_ref_tool, travels to the GPU box), then times
  host decode (this repo's reader, all cores)  |  rcppml_sp_read_gpu (decode + upload, doubles)  |
  Engine.set_matrix_spz (decode + upload + device transpose)  |  the same with the file's transpose section  |
  a k=64 Cholesky fit of 20 iterations on the ingested matrix
and checks the device operands against the decoded arrays.

  python tools/spz_ingest_bench.py [--m 1000000 --n 10000 --per-col 1000] [--out gpurun_out/spz_ingest.json]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from spz_helpers import REF_TOOL  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--n", type=int, default=10_000)
    ap.add_argument("--per-col", type=int, default=1000)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import rcppml_b200 as rb
    m, n, per = args.m, args.n, args.per_col
    rng = np.random.default_rng(1)
    tmp = tempfile.mkdtemp(prefix="spzingest")
    b, s = os.path.join(tmp, "a.bin"), os.path.join(tmp, "a.spz")
    with open(b, "wb") as f:
        np.array([m, n], np.int32).tofile(f)
        np.array([n * per], np.int64).tofile(f)
        (np.arange(n + 1, dtype=np.int64) * per).astype(np.int32).tofile(f)
        for c0 in range(0, n, 5000):
            nb = min(5000, n - c0)
            (np.sort(rng.integers(0, m - per, size=(nb, per), dtype=np.int32), axis=1) + np.arange(per, dtype=np.int32)).tofile(f)
        for c0 in range(0, n, 5000):
            nb = min(5000, n - c0)
            rng.random((nb, per), dtype=np.float32).astype(np.float64).tofile(f)
    t0 = time.perf_counter()
    subprocess.run([REF_TOOL, "encode", b, s, "auto", "0", "1", "2048"], check=True)
    out = {"workload": f"float32 {m}x{n}, {per} non-zeros per column (nnz {n * per}), .spz v2 with transpose section, chunk 2048",
           "file_bytes": os.path.getsize(s), "reference_writer_s": round(time.perf_counter() - t0, 2), "host_cores": os.cpu_count()}
    os.remove(b)
    torch.cuda.set_device(0)

    def best(fn, reps=3):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t)
        return min(ts), r

    with rb.SpzFile(s) as f:
        t, (P, I, X) = best(lambda: f.read(0, None, True, 0, np.float32))
        out["host_decode_A_s"] = round(t, 4)
        t, _ = best(lambda: f.read(1, None, False, 0, np.float32))
        out["host_decode_transpose_section_s"] = round(t, 4)

    def read_gpu():
        g = rb.st_read_gpu(s)
        rb.st_free_gpu(g)
    t, _ = best(read_gpu)
    out["st_read_gpu_s"] = round(t, 4)

    eng = rb.Engine(0)
    t, used = best(lambda: eng.set_matrix_spz(s))
    out["set_matrix_spz_s"] = round(t, 4)
    assert used is False
    p, i, x = eng.get_matrix()
    assert np.array_equal(p, P) and np.array_equal(i, I) and np.array_equal(x, X)
    tp0, ti0, tx0 = eng.get_matrix_t()
    t, used = best(lambda: eng.set_matrix_spz(s, stored_transpose=True))
    out["set_matrix_spz_with_stored_transpose_s"] = round(t, 4)
    assert used is True
    tp, ti, tx = eng.get_matrix_t()
    out["stored_transpose_equals_device_transpose"] = bool(np.array_equal(tp, tp0) and np.array_equal(ti, ti0) and np.array_equal(tx, tx0))
    t, _ = best(lambda: eng.set_matrix(m, n, P, I, X))
    out["set_matrix_from_host_arrays_s"] = round(t, 4)
    eng.init_factors(args.k, 42)
    eng.fit(rb.make_config(args.k, max_iter=3, tol=0.0, solver_mode=1))
    eng.init_factors(args.k, 42)
    t, res = best(lambda: eng.fit(rb.make_config(args.k, max_iter=20, tol=0.0, solver_mode=1)), reps=1)
    out["fit_20_iterations_s"] = round(t, 4)
    out["fit_loop_ms_per_iteration"] = round(res.loop_ms / 20, 4)
    eng.close()
    os.remove(s)
    print(json.dumps(out))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
