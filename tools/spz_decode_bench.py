#!/usr/bin/env python
"""Host decode throughput of the .spz reader (rcppml_b200/csrc/spz_reader.cpp) next to the reference's own
decompress_v2 (oracle/_ref/spz_ref_tool time ...), same files, same machine, same thread counts. CPU only — this is
the host half of the on-disk ingest; needs /root/reference-built oracle/_ref (`make -C oracle ref`).

  python tools/spz_decode_bench.py [--out profiles/r02y_spz_decode.json] [--big] [--c4]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from spz_helpers import REF_TOOL, write_bin  # noqa: E402


def time_mine(path, threads, repeats=7, section=0):
    from rcppml_b200 import streampress as S
    with S.SpzFile(path) as f:
        nc = f.section_cols(section)
        nnz = f.range_nnz(section, 0, nc)
        p = np.zeros(nc + 1, np.int32)
        i = np.zeros(max(nnz, 1), np.int32)
        x = np.zeros(max(nnz, 1), np.float32)
        ip = C.POINTER(C.c_int)
        best = 1e9
        for _ in range(repeats):
            t0 = time.perf_counter()
            rc = f._lib.rcppml_b200_spz_read_f32(f._h, section, 0, nc, 1, threads, p.ctypes.data_as(ip), i.ctypes.data_as(ip),
                                                 x.ctypes.data_as(C.POINTER(C.c_float)))
            best = min(best, time.perf_counter() - t0)
            assert rc == 0
        return best * 1e3, nnz


def time_ref(path, threads, repeats=7):
    out = subprocess.run([REF_TOOL, "time", path, str(repeats), str(threads)], capture_output=True, text=True, check=True)
    return float(out.stdout.strip())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--big", action="store_true", help="add a 1e7-nnz float file (about a minute to write)")
    ap.add_argument("--c4", action="store_true", help="add the C4 shape: 1M x 100K, 1000 float32 non-zeros per column, "
                    "with transpose section (1 GB file, ~3 GB of scratch disk)")
    args = ap.parse_args()
    cores = os.cpu_count()
    rows = []
    tmp = tempfile.mkdtemp(prefix="spzbench")
    cases = [("pbmc3k.spz (reference dataset, uint16 counts, 11 chunks)", os.path.join(ROOT, "oracle", "_ref", "pbmc3k.spz"))]
    rng = np.random.default_rng(3)
    synth = [("float32 100000x20000 density 1e-3, chunk 2048", 100000, 20000, 1e-3, "f"),
             ("uint8 counts 100000x20000 density 1e-3, chunk 2048", 100000, 20000, 1e-3, "u8"),
             ("uint8 counts 20000x20000 density 5e-3 (short gaps), chunk 2048", 20000, 20000, 5e-3, "u8")]
    if args.big:
        synth.append(("float32 1000000x10000 density 1e-3 (C4 / 10), chunk 256", 1000000, 10000, 1e-3, "f"))
    for label, m, n, dens, kind in synth:
        A = sp.random(m, n, density=dens, format="csc", random_state=rng, dtype=np.float64)
        if kind == "u8":
            A.data = np.floor(A.data * 30) + 1
        b, s = os.path.join(tmp, "a.bin"), os.path.join(tmp, f"{len(cases)}.spz")
        write_bin(b, A)
        cc = 256 if m >= 1000000 else 2048
        subprocess.run([REF_TOOL, "encode", b, s, "auto", "0", "0", str(cc)], check=True)
        cases.append((label, s))
    if args.c4:
        m, n, per = 1_000_000, 100_000, 1000
        b, s = os.path.join(tmp, "c4.bin"), os.path.join(tmp, "c4.spz")
        with open(b, "wb") as f:                      # the exchange format of spz_ref_tool, written block-wise
            np.array([m, n], np.int32).tofile(f)
            np.array([n * per], np.int64).tofile(f)
            (np.arange(n + 1, dtype=np.int64) * per).astype(np.int32).tofile(f)
            for _ in range(0, n, 5000):               # strictly increasing rows per column
                (np.sort(rng.integers(0, m - per, size=(5000, per), dtype=np.int32), axis=1) + np.arange(per, dtype=np.int32)).tofile(f)
            for _ in range(0, n, 5000):
                rng.random((5000, per), dtype=np.float32).astype(np.float64).tofile(f)
        subprocess.run([REF_TOOL, "encode", b, s, "auto", "0", "1", "2048"], check=True)
        os.remove(b)
        cases.append(("float32 1000000x100000, 1000 per column (C4: 1e8 non-zeros), chunk 2048, + transpose section", s))
    for label, path in cases:
        row = {"file": label, "file_bytes": os.path.getsize(path)}
        for threads in (1, cores):
            heavy = os.path.getsize(path) > 500e6
            mine, nnz = time_mine(path, threads, 3 if heavy else 7)
            ref = time_ref(path, threads, (1 if threads == 1 else 3) if heavy else 7)
            row[f"threads_{threads}"] = {"reader_ms": round(mine, 2), "reference_ms": round(ref, 2),
                                         "reader_Mnnz_per_s": round(nnz / mine / 1e3, 1),
                                         "reference_Mnnz_per_s": round(nnz / ref / 1e3, 1), "speedup": round(ref / mine, 2)}
        row["nnz"] = int(nnz)
        rows.append(row)
        print(json.dumps(row))
    out = {"what": "host decode of .spz v2 files into CSC (int32 + float32 here; uint32 + double in the reference), best of 7",
           "machine": f"build container, {cores} cores ({open('/proc/cpuinfo').read().split('model name')[1].split(':')[1].splitlines()[0].strip()})",
           "reference": "streampress::v2::decompress_v2 (OpenMP over chunks), -O2", "rows": rows}
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
