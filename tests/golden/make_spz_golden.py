#!/usr/bin/env python
"""Writes tests/golden/spz/*.spz + *.npz: small StreamPress v2 files produced by the REFERENCE's own writer
(streampress::v2::compress_v2, through oracle/_ref/spz_ref_tool — `make -C oracle ref`, needs /root/reference) and the
arrays the REFERENCE's own readers (decompress_v2 with and without reorder, decompress_v2_transpose) return for them.
tests/test_spz_reader.py checks rcppml_b200/csrc/spz_reader.cpp against these without the reference present.

  python tests/golden/make_spz_golden.py
"""
import os
import subprocess
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TOOL = os.path.join(ROOT, "oracle", "_ref", "spz_ref_tool")
OUT = os.path.join(HERE, "spz")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from spz_helpers import read_bin, write_bin  # noqa: E402


def tool(*args):
    subprocess.run([TOOL, *map(str, args)], check=True)


def matrix(rng, m, n, density, kind):
    A = sp.random(m, n, density=density, format="csc", random_state=rng, dtype=np.float64)
    if kind == "u8":
        A.data = np.floor(A.data * 200) + 1
    elif kind == "u16":
        A.data = np.floor(A.data * 60000) + 1
    elif kind == "u32":
        A.data = np.floor(A.data * 4e9) + 1
    else:
        A.data = (A.data - 0.5) * 100
    A.sort_indices()
    return A


# name, m, n, density, kind, precision, row_sort, include_transpose, chunk_cols
CASES = [
    ("u8_t", 60, 45, 0.25, "u8", "auto", 0, 1, 16),
    ("u16_escapes", 3000, 40, 0.004, "u16", "auto", 0, 0, 7),
    ("u32_t", 80, 33, 0.2, "u32", "auto", 0, 1, 8),
    ("f32_t", 200, 100, 0.08, "f", "auto", 0, 1, 32),
    ("f16", 120, 50, 0.1, "f", "fp16", 0, 0, 2048),
    ("quant8_t", 90, 70, 0.15, "f", "quant8", 0, 1, 10),
    ("f64", 64, 39, 0.2, "f", "fp64", 0, 0, 13),
    ("u8_rowsort", 70, 50, 0.2, "u8", "auto", 1, 0, 16),
    ("f32_rowsort_t", 40, 60, 0.5, "f", "fp32", 1, 1, 20),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)
    tmp = os.path.join(OUT, "_tmp.bin")
    for name, m, n, density, kind, precision, row_sort, transpose, chunk_cols in CASES:
        A = matrix(rng, m, n, density, kind)
        if name == "u16_escapes":
            A = A.tolil(); A[:, 5:9] = 0; A = A.tocsc(); A.eliminate_zeros(); A.sort_indices()   # empty columns
        spz = os.path.join(OUT, name + ".spz")
        write_bin(tmp, A)
        tool("encode", tmp, spz, precision, row_sort, transpose, chunk_cols)
        out = {"m": m, "n": n, "a_p": A.indptr.astype(np.int32), "a_i": A.indices.astype(np.int32), "a_x": A.data}
        for tag, reorder in (("r1", 1), ("r0", 0)):
            tool("decode", spz, tmp, reorder)
            _, _, p, i, x = read_bin(tmp)
            out.update({f"{tag}_p": p, f"{tag}_i": i, f"{tag}_x": x})
        if transpose:
            tool("decodet", spz, tmp)
            _, _, p, i, x = read_bin(tmp)
            out.update({"t_p": p, "t_i": i, "t_x": x})
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, os.path.getsize(spz), "bytes")
    os.remove(tmp)


if __name__ == "__main__":
    main()
