"""Column sharding of A across ranks: contiguous column ranges, one per GPU."""
from __future__ import annotations

import numpy as np


def shard_columns(n: int, world: int, rank: int):
    """Even contiguous split of n columns. Returns (first_column, count)."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi - lo


def shard_columns_by_nnz(indptr, world: int):
    """Contiguous column ranges with (nearly) equal non-zero counts — for real matrices whose
    column lengths vary. Returns a list of (first_column, count), one per rank, covering 0..n."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.size - 1
    nnz = int(indptr[-1])
    bounds = [0]
    for r in range(1, world):
        target = (nnz * r) // world
        j = int(np.searchsorted(indptr, target, side="left"))
        bounds.append(min(max(j, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def extract_shard(indptr, indices, data, first: int, count: int):
    """CSC slice A[:, first:first+count] with a zero-based column pointer (no copy of the payload)."""
    indptr = np.asarray(indptr)
    p0, p1 = int(indptr[first]), int(indptr[first + count])
    return (indptr[first:first + count + 1] - p0).astype(np.int32), indices[p0:p1], data[p0:p1]
