"""Column sharding of A across ranks: contiguous column ranges, one per GPU."""
from __future__ import annotations

import numpy as np


def block_of(total: int, world: int, rank: int):
    """The engine's partition (csrc/engine.cu block_of): equal blocks of ceil(total/world) items, the last
    ranks may get fewer (or none). Returns (first, count)."""
    nb = -(-total // world)
    lo = min(total, rank * nb)
    return lo, min(total, lo + nb) - lo


shard_columns = block_of


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


def balanced_cuts(counts, world: int, per_item: float = 0.0):
    """Contiguous ranges balanced by WORK (SURVEY.md §8e "contiguous column ranges balanced by nnz"): item j costs
    counts[j] + per_item (the gather is proportional to its non-zeros, the k x k solve is a constant per column; the
    engine's in-process path uses per_item = k). Returns world+1 ascending cuts, cuts[0] = 0, cuts[world] = len(counts):
    cut r is the first position whose work prefix reaches r/world of the total (csrc/engine.cu balanced_cuts_host)."""
    counts = np.asarray(counts, dtype=np.float64)
    total_items = counts.size
    prefix = np.concatenate([[0.0], np.cumsum(counts + per_item)])
    cuts = [0]
    for r in range(1, world):
        target = prefix[-1] * r / world
        j = int(np.searchsorted(prefix, target, side="left"))
        cuts.append(min(max(j, cuts[-1]), total_items))
    cuts.append(total_items)
    return np.asarray(cuts, dtype=np.int32)


def extract_row_block(indptr, indices, data, first: int, count: int):
    """CSC of A[first:first+count, :] over ALL columns, row ids relative to the block (the W half-step
    operand of the rank owning those rows)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices)
    keep = (indices >= first) & (indices < first + count)
    csum = np.concatenate([[0], np.cumsum(keep)])
    new_ptr = csum[indptr].astype(np.int32)
    return new_ptr, (indices[keep] - first).astype(np.int32), np.asarray(data)[keep]


def extract_shard(indptr, indices, data, first: int, count: int):
    """CSC slice A[:, first:first+count] with a zero-based column pointer (no copy of the payload)."""
    indptr = np.asarray(indptr)
    p0, p1 = int(indptr[first]), int(indptr[first + count])
    return (indptr[first:first + count + 1] - p0).astype(np.int32), indices[p0:p1], data[p0:p1]
