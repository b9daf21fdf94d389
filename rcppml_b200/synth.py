"""Host (numpy) twin of the device synthetic generator (csrc/kernels_sparse.cuh).

For column j: round(m·density) candidate rows SplitMix64::hash(seed, t, j) mod m, sorted and
de-duplicated; value 0.5 + uniform<float>(seed+1, r, j). Bit-identical to the device generator
(all-integer hashing; one fp32 conversion and one fp32 add).
"""
from __future__ import annotations

import numpy as np

_G1 = np.uint64(0x9E3779B97F4A7C15)
_G2 = np.uint64(0x6C62272E07BB0142)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)

SEED_A = 20260101


def splitmix_hash(seed, i, j):
    """rng/rng.hpp:129-138, vectorised over uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        h = np.uint64(seed) + i.astype(np.uint64) * _G1 + j.astype(np.uint64) * _G2
        h = (h ^ (h >> np.uint64(30))) * _M1
        h = (h ^ (h >> np.uint64(27))) * _M2
        return h ^ (h >> np.uint64(31))


def unit_float(h):
    """uniform<float>: float(u64) / float(UINT64_MAX) — float(UINT64_MAX) is 2^64 (rng.hpp:102-104)."""
    return (h.astype(np.float32) / np.float32(18446744073709551616.0)).astype(np.float32)


def synth_csc(m: int, n_local: int, col_begin: int = 0, density: float = 1e-3, seed: int = SEED_A,
              chunk_cols: int = 4096):
    """Returns (indptr int32[n_local+1], indices int32[nnz], data float32[nnz])."""
    cnt = int(round(m * density))
    assert cnt >= 1
    t = np.arange(cnt, dtype=np.uint64)
    ptr = np.zeros(n_local + 1, dtype=np.int64)
    idx_parts, val_parts = [], []
    for c0 in range(0, n_local, chunk_cols):
        c1 = min(n_local, c0 + chunk_cols)
        j = np.arange(col_begin + c0, col_begin + c1, dtype=np.uint64)
        rows = (splitmix_hash(seed, t[None, :], j[:, None]) % np.uint64(m)).astype(np.int64)
        rows.sort(axis=1)
        keep = np.ones_like(rows, dtype=bool)
        keep[:, 1:] = rows[:, 1:] != rows[:, :-1]
        ptr[c0 + 1:c1 + 1] = keep.sum(axis=1)
        jj = np.broadcast_to(j[:, None], rows.shape)[keep]
        r = rows[keep]
        vals = np.float32(0.5) + unit_float(splitmix_hash(seed + 1, r.astype(np.uint64), jj))
        idx_parts.append(r.astype(np.int32))
        val_parts.append(vals.astype(np.float32))
    indptr = np.cumsum(ptr)
    assert indptr[-1] < 2 ** 31
    return indptr.astype(np.int32), np.concatenate(idx_parts), np.concatenate(val_parts)
