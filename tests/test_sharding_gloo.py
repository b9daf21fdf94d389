"""CPU, world_size 2 over gloo: the column-sharded ALS schedule the CUDA engine runs over NCCL
(rcppml_b200/csrc/comm.cu) — local H half-steps, all-reduced row sums and Grams, partial W-update
right-hand sides reduce-scattered by row blocks, row-block solves, all-gather of W_T — restated
with the CPU oracle's primitives and gloo collectives, must reproduce the unsharded oracle fit."""
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


def _sharded_fit(rank, world, A, m, n, k, W0, H0, iters, solver, L1, L2):
    from oracle import oracle as O
    from rcppml_b200 import shard
    lo, cnt = shard.shard_columns(n, world, rank)
    Ap, Ai, Ax = shard.extract_shard(A.indptr, A.indices, A.data, lo, cnt)
    Ai, Ax = np.ascontiguousarray(Ai, np.int32), np.ascontiguousarray(Ax, np.float32)
    Atp, Ati, Atx = O.transpose_csc(Ap, Ai, Ax, m, cnt)
    W_T, H = W0.copy(), H0[lo:lo + cnt].copy()
    m_pad = ((m + world - 1) // world) * world
    mb = m_pad // world
    r0, r1 = rank * mb, min(m, (rank + 1) * mb)
    trAtA = np.float32(_allreduce(np.array([np.sum(Ax.astype(np.float64) ** 2)]))[0])
    hist = []
    G_w = O.gram(W_T)
    for it in range(iters):
        warm = it > 0
        # H half-step on local columns
        G = G_w.copy(); G[np.diag_indices(k)] += np.float32(L2[1])
        O.half_step(Ap, Ai, Ax, W_T, G, H, solver_mode=solver, L1=L1[1], warm_start=warm)
        d = (_allreduce(np.abs(H.astype(np.float64)).sum(axis=0)).astype(np.float32) + np.float32(1e-15))
        H /= d
        G_h = _allreduce(H.astype(np.float64).T @ H.astype(np.float64)).astype(np.float32)
        G_h[np.diag_indices(k)] += np.float32(1e-15)
        # W half-step: partial RHS -> reduce-scatter (all-reduce + slice here) -> row-block solve -> all-gather
        B = np.zeros((m_pad, k), np.float32)
        B[:m] = O.rhs(Atp, Ati, Atx, m, H)
        B = _allreduce(B)[r0:r1].copy()
        B_raw = B.copy()
        G = G_h.copy(); G[np.diag_indices(k)] += np.float32(L2[0])
        Wblk = W_T[r0:r1].copy()
        O.solve_given_rhs(B, G, Wblk, solver_mode=solver, L1=L1[0], warm_start=warm)
        sums = _allreduce(np.concatenate([np.abs(Wblk.astype(np.float64)).sum(axis=0),
                                          [np.sum(Wblk.astype(np.float64) * B_raw.astype(np.float64))]]))
        d = sums[:k].astype(np.float32) + np.float32(1e-15)
        cross = np.float32(sums[k])
        Wblk /= d
        G_w = _allreduce(Wblk.astype(np.float64).T @ Wblk.astype(np.float64)).astype(np.float32)
        G_w[np.diag_indices(k)] += np.float32(1e-15)
        full = np.zeros((m_pad, k), np.float32)
        full[r0:r1] = Wblk
        W_T = _allreduce(full)[:m].copy()                       # all-gather (disjoint blocks)
        recon = np.float32(np.sum((np.outer(d, d) * G_w * G_h).astype(np.float64)))
        hist.append(np.float32(trAtA - np.float32(2) * cross + recon))
    return W_T, H, d, np.array(hist), (lo, cnt)


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
        errs = dict(W=rel_err(W_T, ref.W_T), H=rel_err(H, ref.H[lo:lo + cnt]), d=rel_err(d, ref.d),
                    loss=rel_err(hist, ref.loss_history))
        msgs.append((solver, errs))
        ok = ok and max(errs.values()) <= 1e-5
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
