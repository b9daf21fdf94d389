"""CPU, world_size 2 over gloo: the sharded ALS schedule the CUDA engine runs over NCCL
(rcppml_b200/csrc/engine.cu enqueue_iteration + comm.cu) — column blocks of H solved from A[:,J_g],
row blocks of W solved from A[I_g,:]^T, all-reduced Grams / row sums / cross term, all-gather of the
factor blocks — restated with the CPU oracle's primitives and gloo collectives, must reproduce the
unsharded oracle fit; likewise the sharded explicit-mask schedule (slices of the mask pattern, all-reduced loss) and a
work-balanced partition on the reference's skewed pbmc3k block. Also covers the host-side shard helpers
(rcppml_b200/shard.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import random_csc, rel_err  # noqa: E402


def _allreduce(x):
    t = torch.from_numpy(np.ascontiguousarray(x))
    dist.all_reduce(t)
    return t.numpy()


def _allgather_rows(block, first, total):
    """All-gather of disjoint row blocks (emulated with an all-reduce of zero-padded arrays)."""
    full = np.zeros((total, block.shape[1]), np.float32)
    full[first:first + block.shape[0]] = block
    return _allreduce(full)


def _sharded_fit(rank, world, A, m, n, k, W0, H0, iters, solver, L1, L2, cuts=None, operands=None):
    """cuts = (col_cuts, row_cuts): explicit partition (rcppml_b200.shard.balanced_cuts); None = equal blocks.
    operands = callable(lo, cnt, r0, rc) -> ((Ap, Ai, Ax), (Atp, Ati, Atx)): this rank's A[:, J] and (A[I, :])^T from
    somewhere else than slicing A (the .spz ingest); A may then be None."""
    from oracle import oracle as O
    from rcppml_b200 import shard
    if cuts is None:
        lo, cnt = shard.block_of(n, world, rank)          # my columns of H
        r0, rc = shard.block_of(m, world, rank)           # my rows of W
    else:
        lo, cnt = int(cuts[0][rank]), int(cuts[0][rank + 1] - cuts[0][rank])
        r0, rc = int(cuts[1][rank]), int(cuts[1][rank + 1] - cuts[1][rank])
    if operands is not None:
        (Ap, Ai, Ax), (Atp, Ati, Atx) = operands(lo, cnt, r0, rc)
    else:
        Ap, Ai, Ax = shard.extract_shard(A.indptr, A.indices, A.data, lo, cnt)
        Ai, Ax = np.ascontiguousarray(Ai, np.int32), np.ascontiguousarray(Ax, np.float32)
        Rp, Ri, Rx = shard.extract_row_block(A.indptr, A.indices, A.data, r0, rc)
        Atp, Ati, Atx = O.transpose_csc(Rp, Ri, np.ascontiguousarray(Rx, np.float32), rc, n)   # A[I,:]^T: rc columns
    W_T, H = W0.copy(), H0.copy()                      # replicated
    trAtA = np.float32(_allreduce(np.array([np.sum(Ax.astype(np.float64) ** 2)]))[0])
    hist = []

    def gram64(X):
        return X.astype(np.float64).T @ X.astype(np.float64)

    def finish_gram(S):
        G = _allreduce(S).astype(np.float32)
        G[np.diag_indices(k)] += np.float32(1e-15)
        return G

    G_w = finish_gram(gram64(W_T[r0:r0 + rc]))
    for it in range(iters):
        warm = it > 0
        # H half-step on my columns, gathering from the replicated W_T
        G = G_w.copy(); G[np.diag_indices(k)] += np.float32(L2[1])
        Hb = H[lo:lo + cnt].copy()
        O.half_step(Ap, Ai, Ax, W_T, G, Hb, solver_mode=solver, L1=L1[1], warm_start=warm)
        d = (_allreduce(np.abs(Hb.astype(np.float64)).sum(axis=0)).astype(np.float32) + np.float32(1e-15))
        Hb /= d
        G_h = finish_gram(gram64(Hb))
        H = _allgather_rows(Hb, lo, n)
        # W half-step on my rows, gathering from the replicated H (no right-hand side is exchanged)
        G = G_h.copy(); G[np.diag_indices(k)] += np.float32(L2[0])
        Wb = W_T[r0:r0 + rc].copy()
        B_raw = O.rhs(Atp, Ati, Atx, rc, H)
        O.half_step(Atp, Ati, Atx, H, G, Wb, solver_mode=solver, L1=L1[0], warm_start=warm)
        sums = _allreduce(np.concatenate([np.abs(Wb.astype(np.float64)).sum(axis=0),
                                          [np.sum(Wb.astype(np.float64) * B_raw.astype(np.float64))]]))
        d = sums[:k].astype(np.float32) + np.float32(1e-15)
        cross = np.float32(sums[k])
        Wb /= d
        G_w = finish_gram(gram64(Wb))
        W_T = _allgather_rows(Wb, r0, m)
        recon = np.float32(np.sum((np.outer(d, d) * G_w * G_h).astype(np.float64)))
        hist.append(np.float32(trAtA - np.float32(2) * cross + recon))
    return W_T, H, d, np.array(hist), (lo, cnt)


def _sharded_masked_fit(rank, world, A, M, m, n, k, W0, H0, iters, solver, L1, L2):
    """CPU twin of Engine::enqueue_iteration_masked on `world` ranks: every rank solves its column block of H and its row
    block of W_T from its slices of A AND of the mask pattern (columns J of the pattern, columns I of its transpose);
    the unmodified Grams, the row sums and the explicit loss over the non-masked non-zeros are all-reduced."""
    from oracle import oracle as O
    from rcppml_b200 import shard
    lo, cnt = shard.block_of(n, world, rank)
    r0, rc = shard.block_of(m, world, rank)
    Ap, Ai, Ax = shard.extract_shard(A.indptr, A.indices, A.data, lo, cnt)
    Ai, Ax = np.ascontiguousarray(Ai, np.int32), np.ascontiguousarray(Ax, np.float32)
    Mp, Mi, _ = shard.extract_shard(M.indptr, M.indices, M.data, lo, cnt)
    Mi = np.ascontiguousarray(Mi, np.int32)
    At, Mt = A.T.tocsc(), M.T.tocsc()
    At.sort_indices(); Mt.sort_indices()
    Atp, Ati, Atx = shard.extract_shard(At.indptr, At.indices, At.data, r0, rc)          # (A[I, :])^T: rc columns
    Ati, Atx = np.ascontiguousarray(Ati, np.int32), np.ascontiguousarray(Atx, np.float32)
    Mtp, Mti, _ = shard.extract_shard(Mt.indptr, Mt.indices, Mt.data, r0, rc)
    Mti = np.ascontiguousarray(Mti, np.int32)
    W_T, H = W0.copy(), H0.copy()
    hist = []

    def gram(X):
        G = _allreduce(X.astype(np.float64).T @ X.astype(np.float64)).astype(np.float32)
        G[np.diag_indices(k)] += np.float32(1e-15)
        return G

    G_w = gram(W_T[r0:r0 + rc])
    for it in range(iters):
        warm = it > 0
        Hb = H[lo:lo + cnt].copy()
        O.masked_nnls(Ap, Ai, Ax, m, W_T, G_w, Hb, Mp, Mi, L1=L1[1], L2=L2[1], solver_mode=solver, warm_start=warm)
        d = _allreduce(np.abs(Hb.astype(np.float64)).sum(axis=0)).astype(np.float32) + np.float32(1e-15)
        Hb /= d
        G_h = gram(Hb)
        H = _allgather_rows(Hb, lo, n)
        Wb = W_T[r0:r0 + rc].copy()
        O.masked_nnls(Atp, Ati, Atx, n, H, G_h, Wb, Mtp, Mti, L1=L1[0], L2=L2[0], solver_mode=solver, warm_start=warm)
        d = _allreduce(np.abs(Wb.astype(np.float64)).sum(axis=0)).astype(np.float32) + np.float32(1e-15)
        Wb /= d
        G_w = gram(Wb)
        W_T = _allgather_rows(Wb, r0, m)
        # explicit loss over the non-masked non-zeros of MY columns (masked_nnls.hpp:251-282), then over ranks
        sq = 0.0
        for jl in range(cnt):
            rows = Ai[Ap[jl]:Ap[jl + 1]]
            vals = Ax[Ap[jl]:Ap[jl + 1]]
            keep = ~np.isin(rows, Mi[Mp[jl]:Mp[jl + 1]])
            pred = ((W_T[rows[keep]] * d) * H[lo + jl]).sum(axis=1, dtype=np.float32)
            sq += float(np.sum((vals[keep] - pred).astype(np.float32).astype(np.float64) ** 2))
        hist.append(np.float32(_allreduce(np.array([sq]))[0]))
    return W_T, H, d, np.array(hist)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    ok = True
    msgs = []
    m, n, k, iters = 333, 211, 8, 4
    A = random_csc(m, n, 0.08, 31, ragged=True)
    W0, H0 = O.initialize_factors(k, m, n, 42)
    for solver, L1, L2 in [(0, (0.01, 0.02), (0.0, 0.0)), (1, (0.0, 0.0), (0.01, 0.02))]:
        ref = O.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver,
                        L1=L1, L2=L2, threads=1)
        W_T, H, d, hist, (lo, cnt) = _sharded_fit(rank, world, A, m, n, k, W0, H0, iters, solver, L1, L2)
        errs = dict(W=rel_err(W_T, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d),
                    loss=rel_err(hist, ref.loss_history))
        msgs.append((solver, errs))
        ok = ok and max(errs.values()) <= 1e-5
    # explicit user mask, sharded (Engine::enqueue_iteration_masked): CD and per-column Cholesky
    m, n, k, iters = 260, 170, 6, 3
    A = random_csc(m, n, 0.1, 61, ragged=True)
    M = random_csc(m, n, 0.04, 62)
    M.sort_indices()
    W0, H0 = O.initialize_factors(k, m, n, 42)
    for solver in (0, 1):
        L1, L2 = (0.01, 0.02), (0.02, 0.01)
        ref = O.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver,
                        L1=L1, L2=L2, threads=1, mask=(M.indptr, M.indices))
        W_T, H, d, hist = _sharded_masked_fit(rank, world, A, M, m, n, k, W0, H0, iters, solver, L1, L2)
        errs = dict(W=rel_err(W_T, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d), loss=rel_err(hist, ref.loss_history))
        msgs.append(("masked", solver, errs))
        ok = ok and max(errs.values()) <= 1e-5
    # The reference's own skewed data (pbmc3k block: gene rows with 0 .. 200 non-zeros) under the work-balanced
    # partition the engine's in-process multi-GPU path uses (contiguous ranges cut on the prefix sum of nnz + k).
    from rcppml_b200 import shard
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pbmc3k_500x200.npz"))
    m, n, k, iters = 500, 200, 8, 4
    import scipy.sparse as sp
    A = sp.csc_matrix((g["data"].astype(np.float32), g["indices"], g["indptr"]), shape=(m, n))
    A.sort_indices()
    col_cuts = shard.balanced_cuts(np.diff(A.indptr), world, per_item=k)
    row_cuts = shard.balanced_cuts(np.bincount(A.indices, minlength=m), world, per_item=k)
    W0, H0 = O.initialize_factors(k, m, n, 42)
    ref = O.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=1, threads=1)
    W_T, H, d, hist, _ = _sharded_fit(rank, world, A, m, n, k, W0, H0, iters, 1, (0.0, 0.0), (0.0, 0.0), cuts=(col_cuts, row_cuts))
    errs = dict(W=rel_err(W_T, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d), loss=rel_err(hist, ref.loss_history))
    msgs.append(("pbmc3k balanced", errs, col_cuts.tolist(), row_cuts.tolist()))
    ok = ok and max(errs.values()) <= 1e-5 and row_cuts[1] != m // 2                  # skew: the balanced cut is not the middle
    q.put((rank, ok, msgs))
    dist.destroy_process_group()


def test_sharded_schedule_matches_unsharded_oracle():
    from oracle import oracle as O
    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msgs in results:
        assert ok, (rank, msgs)


def _spz_worker(rank, world, port, q):
    """Sharded ingest of a .spz file, world 2 over gloo: every rank decodes ONLY its column block (a column
    range of the main section) and its row block (a column range of the transpose section, already (A[I, :])^T) with the
    product's reader, runs the sharded schedule on them, and must reproduce the unsharded oracle fit of the whole file."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from rcppml_b200 import shard
    from rcppml_b200 import streampress as S
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spz", "f32_t.spz")
    ok, msgs = True, []
    with S.SpzFile(path) as f:
        m, n = f.shape
        k, iters = 6, 3
        decoded = []

        def operands(lo, cnt, r0, rc):
            a = f.read(0, (lo, lo + cnt), True, 1, np.float32)
            t = f.read(1, (r0, r0 + rc), False, 1, np.float32)
            decoded.append(len(a[1]) + len(t[1]))
            return a, t

        P, I, X = f.read(0)                               # the checker's view of the whole matrix (not the ranks')
        W0, H0 = O.initialize_factors(k, m, n, 42)
        ref = O.nmf_fit(P, I, X, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=1, L1=(0.01, 0.0), threads=1)
        balanced = (shard.balanced_cuts(f.col_counts(0), world, per_item=k), shard.balanced_cuts(f.col_counts(1), world, per_item=k))
        for cuts in (None, balanced):
            W_T, H, d, hist, _ = _sharded_fit(rank, world, None, m, n, k, W0, H0, iters, 1, (0.01, 0.0), (0.0, 0.0), cuts=cuts,
                                              operands=operands)
            errs = dict(W=rel_err(W_T, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d), loss=rel_err(hist, ref.loss_history))
            msgs.append(("spz", cuts is not None, errs, decoded[-1]))
            ok = ok and max(errs.values()) <= 1e-5
            # no rank decoded the whole file: its two blocks hold about 2 / world of the non-zeros of ONE section
            total = torch.tensor([decoded[-1]], dtype=torch.int64)
            dist.all_reduce(total)
            ok = ok and int(total[0]) == 2 * len(I) and decoded[-1] < 1.5 * len(I)
    q.put((rank, ok, msgs))
    dist.destroy_process_group()


def test_sharded_fit_from_an_spz_file_matches_unsharded_oracle():
    from oracle import oracle as O
    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_spz_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msgs in results:
        assert ok, (rank, msgs)
