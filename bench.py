#!/usr/bin/env python
"""bench.py — ALS throughput of the B200 engine on BASELINE.json's synthetic workload.

One "step" = one full ALS iteration (H half-step + W half-step + scalings + loss) of sparse NMF on
the synthetic 1M x 100K, 0.1 %-dense fp32 matrix at k = 64 (BASELINE.json configs[3]).
metric = processed non-zeros per second = nnz x iterations / seconds (iterations/s reported beside it).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--solver cholesky|cd]

* value      : inputs resident in HBM, CUDA events on the engine's stream around exactly K iterations,
               max over ranks.
* e2e        : the same fit through the reference-facing C ABI (rcppml_gpu_nmf_unified_float) from pinned
               HOST buffers: H2D of CSC/W/H (double on the wire), device transpose, K iterations, D2H.
* roofline   : the fused gather+solve kernel (two launches per iteration), algorithmic bytes per launch
                / mean CUDA-event duration of those launches, against MEASURED_PEAKS.json.
* cpu_baseline / --impl reference : the CPU restatement of the reference algorithm (oracle/, OpenMP, all
               host cores) on the same matrix, bounded in wall time. (oracle/_ref/libref_fit.so — the reference's own
               nmf_fit compiled against an Eigen stand-in — reproduces the same factors bit for bit but is a checker,
               8x slower than the port because the stand-in's linear algebra is not optimised; it is not timed.)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_A = 20260101
SEED_INIT = 42
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--solver", default="cholesky", choices=["cholesky", "cd"],
                    help="cholesky = solver_mode 1, what R's nmf() selects for GPU at k>32 (R/nmf_thin.R:368-369)")
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--density", type=float, default=1e-3)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--L1", type=float, default=0.0, help="L1 penalty on both factors (C5: 0.01)")
    ap.add_argument("--L2", type=float, default=0.0, help="L2 penalty on both factors (C5: 0.01)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-blocks", action="store_true",
                    help="N > 1: block-wise factor I/O in the e2e leg (each rank moves only its own blocks over PCIe; "
                         "experiment, off by default until measured)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--also-cd", action="store_true", help="(default now) append the secondary solver_mode=0 measurement")
    ap.add_argument("--no-cd", action="store_true", help="skip the secondary solver_mode=0 (coordinate descent) measurement")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.dev)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(args, world):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/summarize_ncu.py) — only for the workload it was captured on."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    key = f"{args.m}x{args.n}x{args.density:g}_k{args.k}_{args.solver}_n{world}"
    return t.get(key)


def solver_mode(args):
    return 1 if args.solver == "cholesky" else 0


def algorithmic_bytes(nnz, k, m, n, mode):
    """DESIGN.md §5 / SURVEY.md §8d. Per solve launch: CSC stream (8 B/nnz) + one gathered k-vector per
    non-zero (4k B) + the solved factor written (4k B/column) (+ read as warm start in CD mode)."""
    per_launch_h = nnz * 8 + nnz * 4 * k + n * 4 * k * (2 if mode == 0 else 1)
    per_launch_w = nnz * 8 + nnz * 4 * k + m * 4 * k * (2 if mode == 0 else 1)
    b_alg_iter = 2 * nnz * 8 + 2 * nnz * 4 * k + 3 * 4 * k * (m + n)
    return per_launch_h, per_launch_w, b_alg_iter


# ----------------------------------------------------------------------------------------------
def run_cpu_reference(args, steps, warmup, budget_s):
    """The reference algorithm's CPU path (oracle restatement, OpenMP, all host cores), timed on the host.
    Full-size matrix when one iteration fits the budget, else the leading block (same density)."""
    from oracle import oracle as O
    O.build()
    cores = O.max_threads()
    mode = solver_mode(args)
    m, n, k = args.m, args.n, args.k
    # probe on a 1/16 x 1/16 block to pick the sample
    def fit(m_s, n_s, iters, budget):
        Ap, Ai, Ax = O.synth_csc(m, n_s, 0, args.density, SEED_A, m_keep=m_s)
        W0, H0 = O.initialize_factors(k, m_s, n_s, SEED_INIT)
        r = O.nmf_fit(Ap, Ai, Ax, m_s, n_s, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=mode, cd_maxit=100,
                      threads=cores, time_budget_s=budget)
        return r, int(Ap[-1])
    probe, nnz_p = fit(m // 16, n // 16, 2, 0.0)
    t_probe = float(probe.iter_seconds[-1] - probe.iter_seconds[0])           # second iteration (warm)
    est_full = t_probe * 256 * 1.5                                             # nnz x256; cache misses grow
    total_iters = warmup + steps
    per_iter_budget = budget_s / max(1, min(total_iters, 4))
    frac = 1.0
    while est_full * frac * frac > per_iter_budget and frac > 1 / 16:
        frac /= 2
    m_s, n_s = int(m * frac), int(n * frac)
    r, nnz_s = fit(m_s, n_s, total_iters, budget_s * 2)
    its = r.iter_seconds
    w = min(warmup, len(its) - 1)
    timed = len(its) - w
    secs = float(its[-1] - (its[w - 1] if w > 0 else 0.0))
    value = nnz_s * timed / secs
    sample = (f"{'full' if frac == 1.0 else 'leading %dx%d block of the' % (m_s, n_s)} {m}x{n} synthetic matrix "
              f"(nnz {nnz_s}), {timed} timed iterations after {w} warm-up, oracle restatement "
              f"(-O2 -fopenmp -ffp-contract=off), solver_mode={mode}")
    return {"value": value, "unit": "nnz/s", "cores": cores, "kind": "port", "sample": sample,
            "iters_per_sec": timed / secs, "ms_per_step": 1e3 * secs / timed, "steps": timed, "nnz": nnz_s}


def cpu_reference_native_row(args, budget_s):
    """SURVEY.md §8d asks for a second, labelled CPU row: the same restatement built with -O3 -march=native (the
    package itself builds with plain -O2, no -march: src/Makevars:6). Built ON THIS BOX (a -march=native object from
    the build container may not run here) and timed in a child process, so that nothing it does — a failed build,
    an illegal instruction — can take the bench line down. Returns a dict or None."""
    import shutil
    tmp = tempfile.mkdtemp(prefix="oracle_native_")
    try:
        so = os.path.join(tmp, "liboracle_native.so")
        cc = ["/usr/bin/g++", "-std=c++17", "-O3", "-march=native", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared",
              "-o", so, os.path.join(ROOT, "oracle", "nmf_oracle.cpp")]
        if subprocess.run(cc, capture_output=True, timeout=120).returncode != 0:
            return None
        code = ("import json,sys; sys.argv=['bench.py','--m','%d','--n','%d','--density','%r','--k','%d','--solver','%s'];"
                "sys.path.insert(0,%r); import bench; a=bench.parse_args();"
                "r=bench.run_cpu_reference(a, steps=3, warmup=1, budget_s=%r); print('NATIVE_ROW '+json.dumps(r))"
                % (args.m, args.n, args.density, args.k, args.solver, ROOT, budget_s))
        env = dict(os.environ, RCPPML_ORACLE_LIB=so)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
        for ln in out.stdout.splitlines():
            if ln.startswith("NATIVE_ROW "):
                r = json.loads(ln[len("NATIVE_ROW "):])
                return {"value": r["value"], "unit": "nnz/s", "cores": r["cores"], "iters_per_sec": r["iters_per_sec"],
                        "flags": "-O3 -march=native -fopenmp -ffp-contract=off (built on this box)",
                        "sample": r["sample"].replace("-O2 -fopenmp", "-O3 -march=native -fopenmp")}
        return None
    except Exception:
        return None
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def print_reference_line(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu_reference(args, args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": "nnz_per_sec", "value": r["value"], "unit": "nnz/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "iters_per_sec": r["iters_per_sec"],
        "config": {"workload": f"synthetic {args.m}x{args.n} {args.density:g}-dense fp32 CSC, k={args.k}, "
                               f"solver_mode={solver_mode(args)} (CPU sample: {r['sample']})"},
        "cpu_baseline": {"value": r["value"], "unit": "nnz/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def run_e2e(args, eng, steps):
    """Full fit through rcppml_gpu_nmf_unified_float from pinned host buffers (bridge_nmf.hpp packing)."""
    import numpy as np
    import torch
    from rcppml_b200 import bridge

    p, i, x = eng.get_matrix()
    W0, H0, _ = eng.get_factors()

    def pinned(a, dtype):
        t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
        out = t.numpy()
        out[...] = a
        return out, t
    keep = []
    colp, t0 = pinned(p, torch.int32); keep.append(t0)
    rowi, t1 = pinned(i, torch.int32); keep.append(t1)
    vals, t2 = pinned(x.astype(np.float64), torch.float64); keep.append(t2)
    W, t3 = pinned(W0.astype(np.float64), torch.float64); keep.append(t3)
    H, t4 = pinned(H0.astype(np.float64), torch.float64); keep.append(t4)
    # one untimed warm-up call (lazy module load, first big allocations), then the timed call
    Ww, tw = pinned(W0.astype(np.float64), torch.float64); keep.append(tw)
    Hw, th = pinned(H0.astype(np.float64), torch.float64); keep.append(th)
    t_start = time.perf_counter()
    bridge.PackedCall(colp, rowi, vals, args.m, eng.n, args.k, Ww, Hw, max_iter=1, tol=0.0,
                      solver_mode=solver_mode(args), cd_maxit=100)()
    warm_secs = time.perf_counter() - t_start
    call = bridge.PackedCall(colp, rowi, vals, args.m, eng.n, args.k, W, H, max_iter=steps, tol=0.0,
                             solver_mode=solver_mode(args), cd_maxit=100)
    t_start = time.perf_counter()
    call()
    secs = time.perf_counter() - t_start
    assert call.status == 0 and call.iterations == steps, (call.status, call.iterations)
    import ctypes as C
    from rcppml_b200 import _lib
    ph = (C.c_double * 5)()
    _lib.load().rcppml_b200_last_call_phases(ph)
    phases = dict(zip(("matrix_h2d_ms", "transpose_ms", "factors_h2d_ms", "als_loop_ms", "factors_d2h_ms"),
                      (round(float(v), 3) for v in ph)))
    h2d = colp.nbytes + rowi.nbytes + vals.nbytes + W.nbytes + H.nbytes
    d2h = W.nbytes + H.nbytes + 8 * args.k
    return {"value": eng.nnz * steps / secs, "unit": "nnz/s", "h2d_bytes_per_step": h2d // steps,
            "d2h_bytes_per_step": d2h // steps, "seconds_total": secs, "warmup_call_seconds": warm_secs,
            "iters_per_sec": steps / secs, "phases": phases,
            "note": "one rcppml_gpu_nmf_unified_float call from pinned host buffers: H2D (double on the wire) + "
                    f"device transpose + {steps} iterations + D2H; bytes are totals / steps; the engine behind the "
                    "entry point is cached per process, so the timed call reuses the warm-up call's device buffers"}


def run_e2e_sharded(args, eng, dist, steps, rank, world):
    """N > 1: the same fit through the public sharded engine API from pinned HOST buffers on every rank — H2D of
    this rank's column block and row block of A, device transpose, H2D of the initial factors, `steps` iterations,
    D2H of the factors. Wall clock between barriers, max over ranks (the reference ABI has no multi-GPU entry)."""
    import numpy as np
    import scipy.sparse as sp
    import torch
    import rcppml_b200 as rb

    m, n, k = args.m, eng.n, args.k
    cp, ci, cx = eng.get_matrix()                         # A[:, J_g]  (CSC, global row ids)
    tp, ti, tx = eng.get_matrix_t()                       # (A[I_g, :])^T as CSC over the block's rows
    RB = sp.csc_matrix((tx, ti, tp), shape=(n, eng.m_loc)).T.tocsc()     # -> A[I_g, :] (n columns, local row ids)
    RB.sort_indices()
    W0, H0, _ = eng.get_factors()
    keep = []

    def pinned(a, dtype):
        t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
        out = t.numpy()
        out[...] = a
        keep.append(t)
        return out
    cb = (pinned(cp, torch.int32), pinned(ci, torch.int32), pinned(cx, torch.float32))
    rbk = (pinned(RB.indptr.astype(np.int32), torch.int32), pinned(RB.indices.astype(np.int32), torch.int32),
           pinned(RB.data.astype(np.float32), torch.float32))
    W0p, H0p = pinned(W0, torch.float32), pinned(H0, torch.float32)
    # results land in pinned host buffers too (as W / H do in the reference ABI call at N = 1), not in fresh pageable arrays
    blocks = bool(getattr(args, "e2e_blocks", False))
    if blocks:                                            # only this rank's rows of W_T / H cross PCIe, both ways
        W0p = pinned(W0[eng.row_begin:eng.row_begin + eng.m_loc], torch.float32)
        H0p = pinned(H0[eng.col_begin:eng.col_begin + eng.n_loc], torch.float32)
    outs = (pinned(np.zeros_like(W0p), torch.float32), pinned(np.zeros_like(H0p), torch.float32),
            pinned(np.zeros(k, np.float32), torch.float32))
    cfg = rb.make_config(k, max_iter=steps, tol=0.0, solver_mode=solver_mode(args), cd_maxit=100,
                         L1=(args.L1, args.L1), L2=(args.L2, args.L2))

    def one_call(iters):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.set_matrix_sharded(m, n, cb, rbk)
        (eng.set_factor_blocks if blocks else eng.set_factors)(W0p, H0p)
        c = rb.make_config(k, max_iter=iters, tol=0.0, solver_mode=solver_mode(args), cd_maxit=100,
                           L1=(args.L1, args.L1), L2=(args.L2, args.L2))
        res = eng.fit(c)
        out = (eng.get_factor_blocks if blocks else eng.get_factors)(out=outs)
        torch.cuda.synchronize()
        dist.barrier()
        secs = time.perf_counter() - t0
        assert res.status == 0 and res.iterations == iters, res
        return secs, out
    warm_secs, _ = one_call(1)
    secs, _ = one_call(steps)
    t = torch.tensor([secs], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t.item())
    h2d = sum(a.nbytes for a in cb) + sum(a.nbytes for a in rbk) + W0p.nbytes + H0p.nbytes
    d2h = W0p.nbytes + H0p.nbytes + 4 * k
    b = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
    dist.all_reduce(b)
    return {"value": eng.nnz_global * steps / secs, "unit": "nnz/s", "h2d_bytes_per_step": int(b[0].item()) // steps,
            "d2h_bytes_per_step": int(b[1].item()) // steps, "seconds_total": secs, "warmup_call_seconds": warm_secs,
            "iters_per_sec": steps / secs,
            "note": f"sharded engine API on {world} ranks from pinned host buffers: H2D of each rank's column + row "
                    f"block (fp32) and of the initial factors, device transpose, {steps} iterations, D2H of the "
                    "factors on every rank; wall clock between barriers, max over ranks; bytes summed over ranks / steps"
                    + ("; --e2e-blocks: every rank moves only its own factor blocks, replicas completed by an NVLink "
                       "all-gather" if blocks else "")}


def main():
    args = parse_args()
    if args.impl == "reference":
        print_reference_line(args)
        return

    import numpy as np  # noqa: F401
    import torch
    import rcppml_b200 as rb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    mode = solver_mode(args)
    m, n, k = args.m, args.n, args.k
    eng = rb.Engine(local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(rb.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        eng.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    # every rank: its column block A[:,J_g] and row block A[I_g,:] of the same m x n matrix
    eng.set_matrix_synthetic_sharded(m, n, args.density, SEED_A)
    nnz_total = eng.nnz_global

    p2p = False

    def timed_fit(mode_, steps, warmup):
        nonlocal p2p
        eng.init_factors(k, SEED_INIT, 0)
        if dist is not None and not p2p:
            p2p = eng.comm_enable_p2p(dist)                  # NVLink peer-memory loop (RCPPML_B200_P2P=0: NCCL)
        cfg = rb.make_config(k, max_iter=steps + warmup, tol=0.0, solver_mode=mode_, cd_maxit=100,
                             L1=(args.L1, args.L1), L2=(args.L2, args.L2))
        eng.set_profiling(False)
        eng.begin_fit(cfg)
        eng.iterate(warmup)
        launches0 = eng.result().gpu_launches
        _, prof_launch0 = eng.profile()                      # per-section launch counts so far (warm-up iterations)
        eng.set_profiling(True)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        eng.iterate(steps)                                   # CUDA events on the engine stream inside
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        clocks = sampler.stop() if rank == 0 else None
        res = eng.result()
        ms = res.loop_ms
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        prof_ms, prof_launch = eng.profile()
        # section times cover the timed steps only: count only the launches of the timed steps against them
        prof_launch = {kk: v - prof_launch0.get(kk, 0) for kk, v in prof_launch.items()}
        assert res.iterations == steps + warmup and res.status == 0, res
        return ms, res.gpu_launches - launches0, prof_ms, prof_launch, clocks, eng.cd_sweeps()

    ms, launches, prof_ms, prof_launch, clocks, _ = timed_fit(mode, args.steps, args.warmup)
    value = nnz_total * args.steps / (ms / 1e3)

    peak, peak_src = peaks()
    bh, bw, b_iter = algorithmic_bytes(nnz_total, k, m, n, mode)
    solve_ms = prof_ms["fused_rhs_nnls_H"] + prof_ms["fused_rhs_nnls_W"]
    solve_launches = prof_launch["fused_rhs_nnls_H"] + prof_launch["fused_rhs_nnls_W"]
    # per launch, this rank's share of the algorithmic bytes (column shards split nnz evenly)
    per_launch_bytes = (bh + bw) / 2.0 / world
    achieved = per_launch_bytes / (solve_ms / max(1, solve_launches) / 1e3) / 1e9
    traffic = ncu_traffic(args, world)
    roofline = {"bound": "hbm", "kernel": "fused gather + NNLS solve, two launches per iteration: half_step_kernel (H half-step, long columns) and tiled_half_step_kernel (W half-step, short columns)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": (traffic or {}).get("bytes_per_launch_mean"), "traffic_detail": traffic,
                "peak_source": peak_src,
                "per_launch_algorithmic_bytes": per_launch_bytes,
                "mean_launch_ms": solve_ms / max(1, solve_launches),
                "iteration": {"B_alg_bytes": b_iter, "achieved_GBs": b_iter / world / (ms / args.steps / 1e3) / 1e9,
                              "frac": b_iter / world / (ms / args.steps / 1e3) / 1e9 / peak},
                "sections_ms_per_step": {kk: v / args.steps for kk, v in prof_ms.items()}}

    line = {
        "metric": "nnz_per_sec", "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "iters_per_sec": args.steps / (ms / 1e3),
        "config": {"workload": f"synthetic {m}x{n} {args.density:g}-dense fp32 CSC (nnz {nnz_total}), k={k}, "
                               f"ALS iteration = H half-step + W half-step + scaling + loss, tol=0",
                   "solver_mode": mode, "solver": args.solver, "cd_maxit": 100, "L1": args.L1, "L2": args.L2,
                   "seed_A": SEED_A,
                   "seed_init": SEED_INIT,
                   "parallelism": ((f"column blocks of H + row blocks of W over {world} GPUs; solved columns stored "
                                    f"into every replica over NVLink peer memory by the solve kernel, one-shot "
                                    f"peer-memory fp64 all-reduce of Grams/norms (no NCCL call in the loop)")
                                   if p2p else
                                   (f"column blocks of H + row blocks of W over {world} GPUs, NCCL all-gather of the "
                                    f"factor blocks, fp64 all-reduce of Grams/norms")) if world > 1 else "single GPU",
                   "l2_policy": "inputs larger than L2 (CSC 1.6 GB + factors 0.28 GB per iteration vs 126 MB L2); no flush"},
        "clocks": clocks, "gpu_launches": launches, "roofline": roofline,
    }

    # SURVEY.md §8d asks for both solvers: the headline is solver_mode 1 (what R selects for the GPU at k > 32);
    # solver_mode 0 (coordinate descent, cd_maxit 100, cd_tol 1e-8 — the CPU default) is reported beside it.
    if rank == 0 and world == 1 and not args.no_cd and mode != 0:
        st, wu = max(2, args.steps // 4), 3
        ms2, l2, pm2, _, _, sweeps = timed_fit(0, st, wu)
        line["solver_mode_0"] = {"ms_per_step": ms2 / st, "value": nnz_total * st / (ms2 / 1e3), "unit": "nnz/s",
                                 "iters_per_sec": st / (ms2 / 1e3), "steps": st, "warmup": wu,
                                 "cd_sweeps_total_incl_warmup": sweeps, "gpu_launches": l2,
                                 "kernel": "tiled_half_step_kernel<.., SOLVER_CD> (kernels_tiled.cuh + the blocked CD of kernels_cd.cuh)",
                                 "sections_ms_per_step": {kk: v / st for kk, v in pm2.items()}}

    if world == 1 and not args.no_e2e:
        eng.init_factors(k, SEED_INIT, 0)
        line["e2e"] = run_e2e(args, eng, args.steps)
    elif world > 1 and not args.no_e2e:
        eng.init_factors(k, SEED_INIT, 0)
        line["e2e"] = run_e2e_sharded(args, eng, dist, args.steps, rank, world)
    else:
        line["e2e"] = None
    eng.close()

    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = run_cpu_reference(args, steps=3, warmup=1, budget_s=args.cpu_budget_s)
        line["cpu_baseline"] = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["iters_per_sec"] = cpu["iters_per_sec"]
        line["cpu_baseline"]["flags"] = "-O2 -fopenmp -ffp-contract=off (the package's flags: no -march, no FMA)"
        native = cpu_reference_native_row(args, args.cpu_budget_s)
        if native is not None:
            line["cpu_baseline"]["O3_march_native"] = native
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
