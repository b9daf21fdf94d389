#!/usr/bin/env python
"""bench.py — ALS throughput of the B200 engine on BASELINE.json's synthetic workload.

One "step" = one full ALS iteration (H half-step + W half-step + scalings + loss) of sparse NMF on
the synthetic 1M x 100K, 0.1 %-dense fp32 matrix at k = 64 (BASELINE.json configs[3]).
metric = processed non-zeros per second = nnz x iterations / seconds (iterations/s reported beside it).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--solver cholesky|cd]

* value      : inputs resident in HBM, CUDA events on the engine's stream around exactly K iterations,
               max over ranks (min / max over ranks of the loop and of every section are printed too).
* e2e        : the same fit through the reference-facing C ABI (rcppml_gpu_nmf_unified_float) from pinned
               HOST buffers: H2D of CSC/W/H (double on the wire), device transpose, K iterations, D2H.
               N > 1: the SAME entry point called once by rank 0 with RCPPML_NUM_GPUS=N — what the single-threaded
               reference caller does (gpu/bridge_nmf.hpp:187, :310-342); the other ranks wait on the host.
* parity     : N = 1: W, d, H, the loss history and the zero patterns of the engine after the iterations the
               cpu_baseline leg ran, against that oracle run ON THE SAME (full) MATRIX (rel. error, 1e-5 budget);
               N > 1: device checksums of W_T / H / d of the sharded fit equal to those of a one-GPU fit of the same
               seed on every rank (`bit_identical_to_n1`).
* roofline   : the fused gather+solve kernel (two launches per iteration). `achieved` / `frac` keep SURVEY.md
               §8d's ALGORITHMIC bytes (an L2-level quantity: every factor row is gathered ~100x per half-step
               and the re-reads hit the 126 MB L2, so frac > 1 is expected); `frac_dram` = DRAM bytes of the
               committed ncu capture / live launch time / measured HBM peak; `frac_l2` = lts__throughput of the
               same capture. bound = "l2" (DESIGN.md §5).
* cpu_baseline / --impl reference : the CPU restatement of the reference algorithm (oracle/, OpenMP, ALL host
               cores — threads are set explicitly, torchrun's OMP_NUM_THREADS=1 does not apply) on the same matrix.
               (oracle/_ref/libref_fit.so — the reference's own nmf_fit compiled against an Eigen stand-in —
               reproduces the same factors bit for bit but is a checker, 8x slower than the port because the
               stand-in's linear algebra is not optimised; it is not timed.)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_A = 20260101
SEED_INIT = 42
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback
PARITY_TOL = 1e-5             # north_star: W, d, H within 1e-5 relative fp32 tolerance


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--solver", default="cholesky", choices=["cholesky", "cd"],
                    help="cholesky = solver_mode 1, what R's nmf() selects for GPU at k>32 (R/nmf_thin.R:368-369); "
                         "cd = solver_mode 0, the CPU default (R/nmf_thin.R:370-375)")
    # (--rows / --cols / --rank: under torch.distributed.run use these — its parser claims abbreviations such as --m)
    ap.add_argument("--m", "--rows", dest="m", type=int, default=1_000_000)
    ap.add_argument("--n", "--cols", dest="n", type=int, default=100_000)
    ap.add_argument("--density", type=float, default=1e-3)
    ap.add_argument("--k", "--rank", dest="k", type=int, default=64)
    ap.add_argument("--L1", type=float, default=0.0, help="L1 penalty on both factors (C5: 0.01)")
    ap.add_argument("--L2", type=float, default=0.0, help="L2 penalty on both factors (C5: 0.01)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-sharded", action="store_true",
                    help="N > 1: time the e2e leg through the per-rank sharded engine API (round-1 variant) instead of "
                         "the reference ABI call with RCPPML_NUM_GPUS=N")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cd", action="store_true", help="skip the secondary solver_mode=0 (coordinate descent) measurement")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--parity-iters", type=int, default=3, help="N > 1: iterations of the sharded-vs-one-GPU checksum comparison")
    return ap.parse_args(argv)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def host_cores() -> int:
    """All host cores this process may use. torchrun exports OMP_NUM_THREADS=1; the oracle takes its thread count
    as an explicit argument (num_threads clauses), so that variable does not limit it."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


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
    """DRAM bytes per launch (and L2 throughput) of the dominant kernels from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written from tools/summarize_ncu.py digests) — only for the workload it was captured on."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    key = f"{args.m}x{args.n}x{args.density:g}_k{args.k}_{args.solver}_n{world}"
    return t.get(key)


def solver_mode(args):
    return 1 if args.solver == "cholesky" else 0


def workload_string(args):
    """Identical in both arms (the driver compares it): names BASELINE.json's configuration, nothing run-specific."""
    return (f"synthetic {args.m}x{args.n} {args.density:g}-dense fp32 CSC, k={args.k}, solver_mode={solver_mode(args)}, "
            f"L1={args.L1:g}, L2={args.L2:g}; ALS iteration = H half-step + W half-step + scaling + loss, tol=0")


def algorithmic_bytes(nnz, k, m, n, mode):
    """DESIGN.md §5 / SURVEY.md §8d. Per solve launch: CSC stream (8 B/nnz) + one gathered k-vector per
    non-zero (4k B) + the solved factor written (4k B/column) (+ read as warm start in CD mode)."""
    per_launch_h = nnz * 8 + nnz * 4 * k + n * 4 * k * (2 if mode == 0 else 1)
    per_launch_w = nnz * 8 + nnz * 4 * k + m * 4 * k * (2 if mode == 0 else 1)
    b_alg_iter = 2 * nnz * 8 + 2 * nnz * 4 * k + 3 * 4 * k * (m + n)
    b_min_iter = 2 * nnz * 8 + 4 * 4 * k * m + 3 * 4 * k * n          # SURVEY.md §8d compulsory lower bound
    return per_launch_h, per_launch_w, b_alg_iter, b_min_iter


# ----------------------------------------------------------------------------------------------
def run_cpu_reference(args, steps, warmup, budget_s, mode=None, keep_fit=False, prefer_full=False, hard_cap_s=240.0):
    """The reference algorithm's CPU path (oracle restatement, OpenMP, all host cores), timed on the host.
    The full matrix whenever the estimated run fits (always with prefer_full unless it would exceed hard_cap_s),
    else the leading block of the same density. keep_fit: also return the matrix and the fitted factors (parity)."""
    from oracle import oracle as O
    O.build()
    cores = host_cores()
    mode = solver_mode(args) if mode is None else mode
    m, n, k = args.m, args.n, args.k
    L1, L2 = (args.L1, args.L1), (args.L2, args.L2)

    def fit(m_s, n_s, iters, budget):
        Ap, Ai, Ax = O.synth_csc(m, n_s, 0, args.density, SEED_A, m_keep=m_s, threads=cores)
        W0, H0 = O.initialize_factors(k, m_s, n_s, SEED_INIT)
        r = O.nmf_fit(Ap, Ai, Ax, m_s, n_s, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=mode, cd_maxit=100,
                      L1=L1, L2=L2, threads=cores, time_budget_s=budget)
        return r, (Ap, Ai, Ax)
    # probe on a 1/16 x 1/16 block to pick the sample
    probe, _ = fit(max(m // 16, 64), max(n // 16, 64), 2, 0.0)
    t_probe = float(probe.iter_seconds[-1] - probe.iter_seconds[0])           # second iteration (warm)
    est_full = t_probe * 256 * 1.5                                             # nnz x256; cache misses grow
    total_iters = warmup + steps
    frac = 1.0
    if prefer_full:
        while est_full * frac * frac * total_iters > hard_cap_s and frac > 1 / 16:
            frac /= 2
    else:
        per_iter_budget = budget_s / max(1, min(total_iters, 4))
        while est_full * frac * frac > per_iter_budget and frac > 1 / 16:
            frac /= 2
    m_s, n_s = int(m * frac), int(n * frac)
    r, A = fit(m_s, n_s, total_iters, hard_cap_s if prefer_full else budget_s * 2)
    nnz_s = int(A[0][-1])
    its = r.iter_seconds
    w = min(warmup, len(its) - 1)
    timed = len(its) - w
    secs = float(its[-1] - (its[w - 1] if w > 0 else 0.0))
    value = nnz_s * timed / secs
    sample = (f"{'full' if frac == 1.0 else 'leading %dx%d block of the' % (m_s, n_s)} {m}x{n} synthetic matrix "
              f"(nnz {nnz_s}), {timed} timed iterations after {w} warm-up, oracle restatement "
              f"(-O2 -fopenmp -ffp-contract=off), solver_mode={mode}, {cores} threads")
    out = {"value": value, "unit": "nnz/s", "cores": cores, "kind": "port", "sample": sample,
           "iters_per_sec": timed / secs, "ms_per_step": 1e3 * secs / timed, "steps": timed, "nnz": nnz_s,
           "full_matrix": frac == 1.0}
    if keep_fit:
        out["_fit"] = r
        out["_A"] = A
        out["_shape"] = (m_s, n_s)
    return out


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
        code = ("import json,sys; sys.path.insert(0,%r); import bench;"
                "a=bench.parse_args(['--m','%d','--n','%d','--density','%r','--k','%d','--solver','%s','--L1','%r','--L2','%r']);"
                "r=bench.run_cpu_reference(a, steps=3, warmup=1, budget_s=%r); print('NATIVE_ROW '+json.dumps(r))"
                % (ROOT, args.m, args.n, args.density, args.k, args.solver, args.L1, args.L2, budget_s))
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
    """--impl reference: the reference algorithm's CPU path on ALL host cores, on the arm's own configuration. Under
    torchrun (N > 1) rank 0 alone runs and prints; the other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu_reference(args, args.steps, args.warmup, budget_s=120.0, prefer_full=True)
    line = {
        "impl": "reference", "metric": "nnz_per_sec", "value": r["value"], "unit": "nnz/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "iters_per_sec": r["iters_per_sec"],
        "config": {"workload": workload_string(args)},
        "cpu_baseline": {"value": r["value"], "unit": "nnz/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def _pinned(a, dtype, keep):
    import torch
    t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
    out = t.numpy()
    out[...] = a
    keep.append(t)
    return out


def run_e2e(args, csc, W0, H0, steps, n_gpus=1):
    """Full fit through rcppml_gpu_nmf_unified_float from pinned host buffers (bridge_nmf.hpp packing). n_gpus > 1:
    the same single call with RCPPML_NUM_GPUS set — the library shards the work over the devices of this process."""
    import ctypes as C
    import numpy as np
    import torch
    from rcppml_b200 import _lib, bridge

    p, i, x = csc
    n = len(p) - 1
    nnz = int(p[-1])
    keep = []
    colp = _pinned(p, torch.int32, keep)
    rowi = _pinned(i, torch.int32, keep)
    vals = _pinned(x.astype(np.float64), torch.float64, keep)
    W = _pinned(W0.astype(np.float64), torch.float64, keep)
    H = _pinned(H0.astype(np.float64), torch.float64, keep)
    Ww = _pinned(W0.astype(np.float64), torch.float64, keep)
    Hw = _pinned(H0.astype(np.float64), torch.float64, keep)
    old_env = os.environ.get("RCPPML_NUM_GPUS")
    os.environ["RCPPML_NUM_GPUS"] = str(n_gpus)
    try:
        kw = dict(tol=0.0, solver_mode=solver_mode(args), cd_maxit=100, L1=(args.L1, args.L1), L2=(args.L2, args.L2))
        # untimed warm-up calls (lazy module load, contexts on every device, first big allocations), then the timed call
        warm_secs = []
        for _ in range(2 if n_gpus > 1 else 1):
            Ww[...] = W0
            Hw[...] = H0
            t_start = time.perf_counter()
            wc = bridge.PackedCall(colp, rowi, vals, args.m, n, args.k, Ww, Hw, max_iter=1, **kw)
            wc()
            warm_secs.append(time.perf_counter() - t_start)
            assert wc.status == 0, wc.status
        call = bridge.PackedCall(colp, rowi, vals, args.m, n, args.k, W, H, max_iter=steps, **kw)
        t_start = time.perf_counter()
        call()
        secs = time.perf_counter() - t_start
    finally:
        if old_env is None:
            os.environ.pop("RCPPML_NUM_GPUS", None)
        else:
            os.environ["RCPPML_NUM_GPUS"] = old_env
    assert call.status == 0 and call.iterations == steps, (call.status, call.iterations)
    ph = (C.c_double * 5)()
    _lib.load().rcppml_b200_last_call_phases(ph)
    phases = dict(zip(("matrix_h2d_ms", "transpose_ms", "factors_h2d_ms", "als_loop_ms", "factors_d2h_ms"),
                      (round(float(v), 3) for v in ph)))
    phases["library_wall_ms"] = round(float(_lib.load().rcppml_b200_last_call_wall_ms()), 3)
    h2d = colp.nbytes + rowi.nbytes + vals.nbytes + W.nbytes + H.nbytes
    d2h = W.nbytes + H.nbytes + 8 * args.k
    return {"value": nnz * steps / secs, "unit": "nnz/s", "h2d_bytes_per_step": h2d // steps,
            "d2h_bytes_per_step": d2h // steps, "seconds_total": secs, "warmup_call_seconds": warm_secs,
            "iters_per_sec": steps / secs, "phases": phases, "n_gpus": n_gpus, "final_loss": call.train_loss,
            "entry": "rcppml_gpu_nmf_unified_float" + (f" with RCPPML_NUM_GPUS={n_gpus} (one process, one host thread per device)" if n_gpus > 1 else ""),
            "note": "one reference-ABI call from pinned host buffers: H2D (double on the wire) + device transpose + "
                    f"{steps} iterations + D2H; bytes are totals / steps; the engines behind the entry point are cached "
                    "per process, so the timed call reuses the warm-up call's device buffers"
                    + ("; every device uploads only its own column block and factor blocks, row blocks are assembled "
                       "over NVLink" if n_gpus > 1 else "")}, (W, H)


def run_e2e_sharded(args, eng, dist, steps, rank, world):
    """N > 1, round-1 variant (--e2e-sharded): the fit through the per-rank sharded engine API from pinned HOST buffers on
    every rank — H2D of this rank's column block and row block of A, device transpose, H2D of this rank's factor blocks,
    `steps` iterations, D2H of the blocks. Wall clock between barriers, max over ranks."""
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
    cb = (_pinned(cp, torch.int32, keep), _pinned(ci, torch.int32, keep), _pinned(cx, torch.float32, keep))
    rbk = (_pinned(RB.indptr.astype(np.int32), torch.int32, keep), _pinned(RB.indices.astype(np.int32), torch.int32, keep),
           _pinned(RB.data.astype(np.float32), torch.float32, keep))
    W0p = _pinned(W0[eng.row_begin:eng.row_begin + eng.m_loc], torch.float32, keep)
    H0p = _pinned(H0[eng.col_begin:eng.col_begin + eng.n_loc], torch.float32, keep)
    outs = (_pinned(np.zeros_like(W0p), torch.float32, keep), _pinned(np.zeros_like(H0p), torch.float32, keep),
            _pinned(np.zeros(k, np.float32), torch.float32, keep))

    def one_call(iters):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.set_matrix_sharded(m, n, cb, rbk)
        eng.set_factor_blocks(W0p, H0p)
        c = rb.make_config(k, max_iter=iters, tol=0.0, solver_mode=solver_mode(args), cd_maxit=100,
                           L1=(args.L1, args.L1), L2=(args.L2, args.L2))
        res = eng.fit(c)
        out = eng.get_factor_blocks(out=outs)
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
            "iters_per_sec": steps / secs, "entry": "rcppml_b200_* sharded engine API, one process per GPU",
            "note": f"sharded engine API on {world} ranks from pinned host buffers (fp32 on the wire): H2D of each rank's "
                    f"column + row block and factor blocks, device transpose, {steps} iterations, D2H of the blocks; "
                    "wall clock between barriers, max over ranks; bytes summed over ranks / steps"}


# ----------------------------------------------------------------------------------------------
def rel_err(a, b):
    import numpy as np
    den = float(np.abs(b).max())
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(den, 1e-30))


def parity_vs_oracle(args, eng, cpu, mode):
    """N = 1. The engine against the oracle run the cpu_baseline leg just made: same matrix (the full benchmark matrix
    when that run used it — then the engine's own device-generated matrix, the object that was timed), same initial
    factors, same number of iterations. Every row of W and H is compared (>= 10 000 asked)."""
    import numpy as np
    import rcppml_b200 as rb
    from oracle import oracle as O

    ref = cpu["_fit"]
    m_s, n_s = cpu["_shape"]
    iters = int(ref.iterations)
    k = args.k
    own = None
    if cpu["full_matrix"]:
        e = eng
        e.init_factors(k, SEED_INIT, 0)
        gp, gi, gx = e.get_matrix()
        Ap, Ai, Ax = cpu["_A"]
        same_matrix = bool(np.array_equal(gp, Ap) and np.array_equal(gi, Ai) and np.array_equal(gx, Ax))
    else:
        own = e = rb.Engine(0)
        Ap, Ai, Ax = cpu["_A"]
        e.set_matrix(m_s, n_s, Ap, Ai, Ax)
        W0, H0 = O.initialize_factors(k, m_s, n_s, SEED_INIT)
        e.set_factors(W0, H0)
        same_matrix = True
    cfg = rb.make_config(k, max_iter=iters, tol=0.0, solver_mode=mode, cd_maxit=100, L1=(args.L1, args.L1),
                         L2=(args.L2, args.L2))
    res = e.fit(cfg)
    W, H, d = e.get_factors()
    hist = e.loss_history(iters)
    sweeps = e.cd_sweeps()
    if own is not None:
        own.close()
    errs = {"W": rel_err(W, ref.W_T), "H": rel_err(H, ref.H), "d": rel_err(d, ref.d),
            "loss_history": rel_err(hist, ref.loss_history)}
    zero_w = bool(np.array_equal(W == 0, ref.W_T == 0))
    zero_h = bool(np.array_equal(H == 0, ref.H == 0))
    out = {"against": "oracle (oracle/nmf_oracle.cpp, the cpu_baseline run)", "matrix": cpu["sample"].split(",")[0],
           "oracle_pinning": "the oracle reproduces the reference's own nmf_fit (nmf/fit_cpu.hpp compiled unmodified against an "
                             "Eigen stand-in) bit for bit (tests/test_reference_fit.py); the rounding INSIDE Eigen — notably its "
                             "fp32 rankUpdate for the Gram, here fp64-accumulated and rounded once — is the stand-in's "
                             "definition and cannot be pinned without Eigen (DESIGN.md section 3)",
           "matrix_identical_to_oracle_generator": same_matrix, "iterations": iters, "rows_compared_W": int(W.shape[0]),
           "rows_compared_H": int(H.shape[0]), "rel_err": errs, "max_rel_err": max(errs.values()),
           "zero_pattern_equal": zero_w and zero_h, "bit_identical_W": bool(np.array_equal(W, ref.W_T)),
           "bit_identical_H": bool(np.array_equal(H, ref.H)), "tolerance": PARITY_TOL,
           "ok": bool(res.status == 0 and res.iterations == iters and max(errs.values()) <= PARITY_TOL and zero_w and zero_h
                      and same_matrix)}
    if mode == 0:
        out["cd_sweeps"] = {"engine": int(sweeps), "oracle": int(ref.cd_sweeps), "equal": int(sweeps) == int(ref.cd_sweeps)}
    return out


def parity_vs_one_gpu(args, eng, dist, local_rank, mode, iters, want_matrix=False):
    """N > 1. Every rank fits the WHOLE matrix on its own GPU with a one-GPU engine (same seed, `iters` iterations) and
    compares device checksums of W_T / H / d (and the loss history) with the sharded fit: equal <=> bit-identical."""
    import numpy as np
    import torch
    import rcppml_b200 as rb

    k = args.k
    cfg = rb.make_config(k, max_iter=iters, tol=0.0, solver_mode=mode, cd_maxit=100, L1=(args.L1, args.L1),
                         L2=(args.L2, args.L2))
    eng.init_factors(k, SEED_INIT, 0)
    eng.comm_enable_p2p(dist)
    res = eng.fit(cfg)
    cs = eng.factor_checksum()
    hist = eng.loss_history(iters)
    one = rb.Engine(local_rank)
    one.set_matrix_synthetic(args.m, args.n, 0, args.density, SEED_A)
    one.init_factors(k, SEED_INIT, 0)
    res1 = one.fit(cfg)
    cs1 = one.factor_checksum()
    hist1 = one.loss_history(iters)
    extra = None
    if want_matrix:
        extra = (one.get_matrix(), one.nnz)
    one.close()
    same = bool(cs == cs1 and res.status == 0 and res1.status == 0 and res.iterations == iters)
    loss_err = rel_err(hist, hist1)
    t = torch.tensor([1.0 if same else 0.0, -loss_err], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"against": "one-GPU fit of the same seed on every rank (device checksums of W_T, H, d)",
            "iterations": iters, "bit_identical_to_n1": bool(t[0].item() == 1.0),
            "loss_history_rel_err_max_over_ranks": float(-t[1].item()), "tolerance": PARITY_TOL,
            "ok": bool(t[0].item() == 1.0 and -t[1].item() <= PARITY_TOL)}, extra


def main():
    args = parse_args()
    if args.impl == "reference":
        print_reference_line(args)
        return

    import numpy as np
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
    host_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")      # host-side waits (no spinning NCCL kernel on an idle rank's GPU)

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

    def over_ranks(vals):
        """min / max over ranks of a list of floats."""
        if dist is None:
            return list(vals), list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return [float(v) for v in lo.tolist()], [float(v) for v in hi.tolist()]

    def one_fit(mode_, steps, warmup, profiled):
        """warmup untimed iterations, then exactly `steps` iterations between barrier + synchronize on both sides, timed with
        CUDA events on the engine's stream. profiled=False: the product configuration (no per-section events; the
        steady-state iteration is replayed as a CUDA graph) — the run `value` comes from. profiled=True: the same fit
        once more with per-section CUDA events (plain launches), for the section / per-kernel times only."""
        nonlocal p2p
        eng.init_factors(k, SEED_INIT, 0)
        if dist is not None:
            p2p = eng.comm_enable_p2p(dist)                  # NVLink peer-memory loop (RCPPML_B200_P2P=0: NCCL)
        cfg = rb.make_config(k, max_iter=steps + warmup, tol=0.0, solver_mode=mode_, cd_maxit=100,
                             L1=(args.L1, args.L1), L2=(args.L2, args.L2))
        eng.set_profiling(False)
        eng.begin_fit(cfg)
        eng.iterate(warmup)
        launches0 = eng.result().gpu_launches
        _, prof_launch0 = eng.profile()                      # per-section launch counts so far (warm-up iterations)
        eng.set_profiling(profiled)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0 and not profiled:
            sampler.start()
        eng.iterate(steps)                                   # CUDA events on the engine stream inside
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        clocks = sampler.stop() if (rank == 0 and not profiled) else None
        res = eng.result()
        prof_ms, prof_launch = eng.profile()
        eng.set_profiling(False)
        assert res.iterations == steps + warmup and res.status == 0, res
        # section times cover the timed steps only: count only the launches of the timed steps against them
        prof_launch = {kk: v - prof_launch0.get(kk, 0) for kk, v in prof_launch.items()}
        return res, res.gpu_launches - launches0, prof_ms, prof_launch, clocks

    def timed_fit(mode_, steps, warmup):
        res, launches_, _, _, clocks_ = one_fit(mode_, steps, warmup, profiled=False)
        sweeps_ = eng.cd_sweeps()
        res_p, _, prof_ms_, prof_launch_, _ = one_fit(mode_, steps, warmup, profiled=True)
        names = list(prof_ms_.keys())
        lo, hi = over_ranks([res.loop_ms, res_p.loop_ms] + [prof_ms_[s_] for s_ in names])
        ms_ = hi[0]                                          # max over ranks of the product-configuration loop
        spread_ = {"loop_ms_per_step": {"min": lo[0] / steps, "max": hi[0] / steps},
                   "profiled_loop_ms_per_step": {"min": lo[1] / steps, "max": hi[1] / steps},
                   "sections_ms_per_step": {s_: {"min": lo[2 + i] / steps, "max": hi[2 + i] / steps} for i, s_ in enumerate(names)}}
        return ms_, launches_, prof_ms_, prof_launch_, clocks_, sweeps_, spread_

    ms, launches, prof_ms, prof_launch, clocks, _, spread = timed_fit(mode, args.steps, args.warmup)
    value = nnz_total * args.steps / (ms / 1e3)

    peak, peak_src = peaks()
    bh, bw, b_iter, b_min = algorithmic_bytes(nnz_total, k, m, n, mode)
    solve_ms = prof_ms["fused_rhs_nnls_H"] + prof_ms["fused_rhs_nnls_W"]
    solve_launches = prof_launch["fused_rhs_nnls_H"] + prof_launch["fused_rhs_nnls_W"]
    # per launch, this rank's share of the algorithmic bytes (column shards split nnz evenly)
    per_launch_bytes = (bh + bw) / 2.0 / world
    mean_launch_ms = solve_ms / max(1, solve_launches)
    achieved = per_launch_bytes / (mean_launch_ms / 1e3) / 1e9
    traffic = ncu_traffic(args, world)
    dram_per_launch = (traffic or {}).get("bytes_per_launch_mean")
    roofline = {"bound": "l2" if mode == 1 else "issue",
                "bound_note": ("the algorithmic bytes arrive from the 126 MB L2, not from HBM (each factor row is gathered ~100x "
                               "per half-step): `frac` is on SURVEY.md §8d's no-reuse B_alg scale and may exceed 1; `frac_dram` "
                               "is the HBM fraction, `frac_l2` the L2 -> SM throughput fraction of the committed ncu capture")
                              if mode == 1 else "coordinate descent is instruction-issue bound (DESIGN.md §4); HBM fractions are not the limiter",
                "kernel": "fused gather + NNLS solve, two launches per iteration (H half-step over the columns of A, W half-step over the rows)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": dram_per_launch, "traffic_detail": traffic,
                "frac_dram": (dram_per_launch / (mean_launch_ms / 1e3) / 1e9 / peak) if dram_per_launch else None,
                "frac_l2": (traffic or {}).get("lts_throughput_frac_mean"),
                "peak_source": peak_src,
                "per_launch_algorithmic_bytes": per_launch_bytes,
                "mean_launch_ms": mean_launch_ms,
                "iteration": {"B_alg_bytes": b_iter, "B_min_bytes": b_min,
                              "achieved_GBs": b_iter / world / (ms / args.steps / 1e3) / 1e9,
                              "frac": b_iter / world / (ms / args.steps / 1e3) / 1e9 / peak,
                              "frac_B_min": b_min / world / (ms / args.steps / 1e3) / 1e9 / peak},
                "sections_ms_per_step": {kk: v / args.steps for kk, v in prof_ms.items()},
                "over_ranks": spread}

    line = {
        "metric": "nnz_per_sec", "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "iters_per_sec": args.steps / (ms / 1e3),
        "config": {"workload": workload_string(args), "nnz": nnz_total,
                   "solver_mode": mode, "solver": args.solver, "cd_maxit": 100, "L1": args.L1, "L2": args.L2,
                   "seed_A": SEED_A,
                   "seed_init": SEED_INIT,
                   "parallelism": ((f"column blocks of H + row blocks of W over {world} GPUs; solved columns stored into "
                                    f"every replica by the solve kernel with NVSwitch multicast stores (multimem.st: one "
                                    f"store per word, the switch replicates), one-shot peer-memory fp64 all-reduce of "
                                    f"Grams/norms (no NCCL call in the loop)")
                                   if (p2p and getattr(eng, "p2p_mode", "") == "multicast") else
                                   (f"column blocks of H + row blocks of W over {world} GPUs; solved columns stored "
                                    f"into every replica over NVLink peer memory by the solve kernel, one-shot "
                                    f"peer-memory fp64 all-reduce of Grams/norms (no NCCL call in the loop)")
                                   if p2p else
                                   (f"column blocks of H + row blocks of W over {world} GPUs, NCCL all-gather of the "
                                    f"factor blocks, fp64 all-reduce of Grams/norms")) if world > 1 else "single GPU",
                   "l2_policy": "inputs larger than L2 (CSC 1.6 GB + factors 0.28 GB per iteration vs 126 MB L2); no flush"},
        "clocks": clocks, "gpu_launches": launches, "roofline": roofline,
    }

    # SURVEY.md §8d asks for both solvers: the headline is solver_mode 1 (what R selects for the GPU at k > 32);
    # solver_mode 0 (coordinate descent, cd_maxit 100, cd_tol 1e-8 — the CPU default) is reported beside it, with the
    # CPU restatement in the same mode on a bounded sample so that a CD ratio exists too.
    if rank == 0 and world == 1 and not args.no_cd and mode != 0:
        st, wu = max(2, args.steps // 4), 3
        ms2, l2, pm2, _, _, sweeps, _ = timed_fit(0, st, wu)
        line["solver_mode_0"] = {"ms_per_step": ms2 / st, "value": nnz_total * st / (ms2 / 1e3), "unit": "nnz/s",
                                 "iters_per_sec": st / (ms2 / 1e3), "steps": st, "warmup": wu,
                                 "cd_sweeps_total_incl_warmup": sweeps, "gpu_launches": l2,
                                 "kernel": "tiled_half_step_kernel<.., SOLVER_CD> (kernels_tiled.cuh + the blocked CD of kernels_cd.cuh)",
                                 "sections_ms_per_step": {kk: v / st for kk, v in pm2.items()}}
        if not args.no_cpu:
            cd_cpu = run_cpu_reference(args, steps=2, warmup=1, budget_s=args.cpu_budget_s, mode=0)
            line["solver_mode_0"]["cpu_reference"] = {kk: cd_cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample", "iters_per_sec")}
            line["solver_mode_0"]["vs_cpu_reference"] = line["solver_mode_0"]["value"] / cd_cpu["value"]

    # ---- parity on this very run's configuration
    parity = None
    csc_host = None
    cpu = None
    if world == 1:
        if not args.no_cpu:
            # the FULL benchmark matrix (1 warm-up + 3 timed iterations: ~7 s of 16 cores at C4) unless that would take
            # longer than 90 s — parity is then checked on the object that was timed
            cpu = run_cpu_reference(args, steps=3, warmup=1, budget_s=args.cpu_budget_s, keep_fit=not args.no_parity,
                                    prefer_full=True, hard_cap_s=90.0)
            if not args.no_parity:
                parity = parity_vs_oracle(args, eng, cpu, mode)
    elif not args.no_parity:
        parity, extra = parity_vs_one_gpu(args, eng, dist, local_rank, mode, args.parity_iters,
                                          want_matrix=(rank == 0 and not args.no_e2e and not args.e2e_sharded))
        if extra is not None:
            csc_host = extra[0]
    line["parity"] = parity

    # ---- end to end through the reference ABI
    line["e2e"] = None
    if not args.no_e2e:
        if world == 1:
            eng.init_factors(k, SEED_INIT, 0)
            W0, H0, _ = eng.get_factors()
            line["e2e"], _ = run_e2e(args, eng.get_matrix(), W0, H0, args.steps, 1)
        elif args.e2e_sharded:
            eng.init_factors(k, SEED_INIT, 0)
            line["e2e"] = run_e2e_sharded(args, eng, dist, args.steps, rank, world)
        else:
            # One process, one call: rank 0 drives all N devices through the reference entry point; the other ranks
            # wait on the HOST (gloo) so that nothing of theirs runs on the GPUs meanwhile.
            eng.init_factors(k, SEED_INIT, 0)
            W0 = H0 = None
            if rank == 0:
                W0, H0, _ = eng.get_factors()
                if csc_host is None:
                    one = rb.Engine(local_rank)
                    one.set_matrix_synthetic(m, n, 0, args.density, SEED_A)
                    csc_host = one.get_matrix()
                    one.close()
            torch.cuda.synchronize()
            dist.barrier(group=host_group)
            if rank == 0:
                e2e, (We, He) = run_e2e(args, csc_host, W0, H0, args.steps, world)
                # the call's factors against the one-process-per-GPU engine after the same number of iterations
                cfg = rb.make_config(k, max_iter=args.steps, tol=0.0, solver_mode=mode, cd_maxit=100,
                                     L1=(args.L1, args.L1), L2=(args.L2, args.L2))
                e2e["_cmp"] = (We, He, cfg)
                line["e2e"] = e2e
            dist.barrier(group=host_group)
            # every rank: the sharded engine's fit of the same length, to compare the ABI call's factors with
            cfg = rb.make_config(k, max_iter=args.steps, tol=0.0, solver_mode=mode, cd_maxit=100,
                                 L1=(args.L1, args.L1), L2=(args.L2, args.L2))
            eng.comm_enable_p2p(dist)
            eng.fit(cfg)
            if rank == 0:
                Ws, Hs, _ = eng.get_factors()
                We, He, _ = line["e2e"].pop("_cmp")
                line["e2e"]["factors_bit_identical_to_sharded_engine"] = bool(
                    np.array_equal(We.astype(np.float32), Ws) and np.array_equal(He.astype(np.float32), Hs))
    eng.close()

    if rank == 0 and world == 1 and cpu is not None:
        line["cpu_baseline"] = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["iters_per_sec"] = cpu["iters_per_sec"]
        line["cpu_baseline"]["flags"] = "-O2 -fopenmp -ffp-contract=off (the package's flags: no -march, no FMA)"
        native = cpu_reference_native_row(args, args.cpu_budget_s)
        if native is not None:
            line["cpu_baseline"]["O3_march_native"] = native
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(group=host_group)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
