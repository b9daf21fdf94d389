"""Raw exchange format of oracle/ref_fixture/spz_ref_tool.cpp ("csc.bin"): int32 m, n; int64 nnz; int32 p[n+1];
int32 i[nnz]; float64 x[nnz]."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "spz_ref_tool")


def write_bin(path, A):
    A = A.tocsc()
    A.sort_indices()
    with open(path, "wb") as f:
        np.array([A.shape[0], A.shape[1]], np.int32).tofile(f)
        np.array([A.nnz], np.int64).tofile(f)
        A.indptr.astype(np.int32).tofile(f)
        A.indices.astype(np.int32).tofile(f)
        A.data.astype(np.float64).tofile(f)


def read_bin(path):
    with open(path, "rb") as f:
        m, n = np.fromfile(f, np.int32, 2)
        nnz = int(np.fromfile(f, np.int64, 1)[0])
        p = np.fromfile(f, np.int32, n + 1)
        i = np.fromfile(f, np.int32, nnz)
        x = np.fromfile(f, np.float64, nnz)
    return int(m), int(n), p, i, x


def ref_tool(*args):
    """Runs the reference codec tool; returns (returncode, stderr)."""
    r = subprocess.run([REF_TOOL, *map(str, args)], capture_output=True, text=True)
    return r.returncode, r.stderr
