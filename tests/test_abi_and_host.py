"""CPU: the C-ABI library loads and exports every symbol include/rcppml_gpu.h declares (no compute
without a GPU), fails loudly instead of falling back, and the host-side logic (packing, sharding,
synthetic generator twin) is right."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rcppml_gpu.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rcppml_(?:gpu|b200)_\w+)\s*\(", src)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from rcppml_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert "rcppml_gpu_nmf_unified_float" in names and "rcppml_gpu_detect" in names and len(names) >= 20
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/rcppml_gpu.h but not exported"


def test_library_is_sm100a_only():
    from rcppml_b200 import _lib
    out = os.popen(f"cuobjdump -lelf {_lib.LIB_PATH} 2>/dev/null").read()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    import rcppml_b200 as rb
    from rcppml_b200._lib import NativeLibraryError
    info = rb.gpu_detect()
    assert info["num_gpus"] == 0 and info["status"] == -1             # src/gpu_bridge_cluster.cu:33-36
    with pytest.raises(NativeLibraryError):
        rb.Engine(0)
    W0 = np.random.default_rng(0).random((30, 4))
    H0 = np.random.default_rng(1).random((20, 4))
    indptr = np.arange(21, dtype=np.int32)
    out = rb.bridge_nmf_sparse(indptr, np.arange(20, dtype=np.int32) % 30, np.ones(20), 30, 20, 4, W0, H0, max_iter=2)
    assert out.status == -1 and out.iterations == 0                   # caller (nmf/fit.hpp:125-133) owns the CPU path
    assert np.array_equal(out.W_T, W0.astype(np.float32))


def test_bridge_packs_73_pointers_in_reference_order():
    from rcppml_b200.bridge import PackedCall
    m, n, k = 6, 5, 3
    indptr = np.array([0, 1, 2, 3, 4, 5], np.int32)
    call = PackedCall(indptr, np.zeros(5, np.int32), np.ones(5), m, n, k, np.ones((m, k)), np.ones((n, k)),
                      max_iter=7, tol=0.5, L1=(0.1, 0.2), L2=(0.3, 0.4), upper_bound=(1.0, 2.0), nonneg=(True, False),
                      cd_maxit=11, seed=9, patience=4, norm_type=1, solver_mode=1)
    a = call._args
    assert len(a) == 73                                               # gpu/bridge_nmf.hpp:39-75
    val = lambda x: x._obj.value
    assert [val(a[i]) for i in (3, 4, 5, 6)] == [m, n, 5, k]
    assert val(a[10]) == 7 and val(a[11]) == 0.5
    assert [val(a[i]) for i in (12, 13, 14, 15)] == [0.2, 0.1, 0.4, 0.3]   # L1_H, L1_W, L2_H, L2_W
    assert [val(a[i]) for i in (20, 21)] == [2.0, 1.0]                     # ub_H, ub_W
    assert val(a[22]) == 11 and val(a[24]) == 9 and val(a[26]) == 4
    assert [val(a[i]) for i in (27, 28)] == [1, 0]                         # nonneg_W, nonneg_H
    assert val(a[33]) == 1 and val(a[36]) == 1                             # norm_type, solver_mode


def test_config_mapping():
    import rcppml_b200 as rb
    c = rb.make_config(8, L1=(0.1, 0.2), L2=(0.3, 0.4), upper_bound=(1, 2), nonneg=(False, True), solver_mode=1)
    assert (c.L1_W, c.L1_H) == (np.float32(0.1), np.float32(0.2)) and (c.ub_W, c.ub_H) == (1.0, 2.0)
    assert (c.nonneg_W, c.nonneg_H) == (0, 1) and c.cd_maxit == 100 and c.patience == 5


def test_shard_helpers():
    from rcppml_b200 import shard
    for n, world in [(100000, 8), (9001, 2), (7, 8), (10, 3)]:
        parts = [shard.block_of(n, world, r) for r in range(world)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        nb = -(-n // world)
        assert all(lo == min(n, r * nb) and c <= nb for r, (lo, c) in enumerate(parts))   # equal all-gather blocks
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 50, 1000)
    indptr = np.concatenate([[0], np.cumsum(counts)])
    parts = shard.shard_columns_by_nnz(indptr, 4)
    assert parts[0][0] == 0 and sum(c for _, c in parts) == 1000
    loads = [indptr[lo + c] - indptr[lo] for lo, c in parts]
    assert max(loads) - min(loads) <= 2 * counts.max()
    p, i, x = shard.extract_shard(indptr, np.arange(indptr[-1]), np.arange(indptr[-1], dtype=np.float32), 10, 5)
    assert p[0] == 0 and p[-1] == i.size == x.size == indptr[15] - indptr[10]


def test_synth_numpy_twin_matches_oracle_generator(oracle):
    from rcppml_b200 import synth
    for (m, n_local, col_begin, dens) in [(5000, 300, 0, 0.01), (20000, 64, 77, 0.05)]:
        a = synth.synth_csc(m, n_local, col_begin, dens)
        b = oracle.synth_csc(m, n_local, col_begin, dens)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    p, i, x = synth.synth_csc(100000, 50, 0, 1e-3)
    assert np.all(np.diff(p) <= 100) and np.all(np.diff(p) >= 95)       # ~0.1 % dedupe loss
    assert x.min() >= 0.5 and x.max() < 1.5 and np.all(i < 100000)
    for j in range(50):
        assert np.all(np.diff(i[p[j]:p[j + 1]]) > 0)                   # sorted, unique rows


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a small matrix: one JSON line
    with the contract's keys, cpu_baseline describing the run, e2e repeating the line's own value with no copies."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--m", "20000", "--n", "2000",
                          "--density", "0.005", "--k", "16", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "nnz/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_bench_sharded_e2e_host_logic(monkeypatch):
    """bench.run_e2e_sharded (the N > 1 end-to-end leg) with a stand-in engine and process group: the host code path the
    driver's scaling run takes — pinned staging of the blocks, result buffers handed to get_factors(out=...), byte
    accounting (--e2e-sharded variant, block-wise factor I/O). No CUDA: pinned allocation and device tensors are mapped
    to plain host ones for the duration of the test."""
    import sys
    import types
    import scipy.sparse as sp
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench
    real_empty, real_tensor = torch.empty, torch.tensor
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "pin_memory"}))
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != "device"}))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    m, n, k = 60, 40, 4
    A = sp.random(m, n, density=0.2, format="csc", dtype=np.float32, random_state=np.random.default_rng(0))
    A.sort_indices()
    T = A[:30, :].tocsc().T.tocsc()
    T.sort_indices()
    calls = []

    class Eng:
        n, m, m_loc, n_loc, row_begin, col_begin, nnz_global = 40, 60, 30, 20, 0, 20, A.nnz

        def get_matrix(self):
            return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data

        def get_matrix_t(self):
            return T.indptr.astype(np.int32), T.indices.astype(np.int32), T.data

        def get_factors(self, out=None):
            if out is None:
                return np.ones((m, k), np.float32), np.ones((n, k), np.float32), np.ones(k, np.float32)
            assert [o.shape for o in out] == [(m, k), (n, k), (k,)] and all(o.dtype == np.float32 for o in out)
            calls.append("get_factors")
            return out

        def get_factor_blocks(self, out=None):
            assert [o.shape for o in out] == [(30, k), (20, k), (k,)]
            calls.append("get_factor_blocks")
            return out

        def set_matrix_sharded(self, m_, n_, cb, rbk):
            assert len(cb) == 3 and len(rbk) == 3 and rbk[0].shape == (n + 1,)

        def set_factors(self, W, H):
            calls.append(("set_factors", W.shape, H.shape))

        def set_factor_blocks(self, W, H):
            calls.append(("set_factor_blocks", W.shape, H.shape))

        def fit(self, c):
            return types.SimpleNamespace(status=0, iterations=c.max_iter)

    class Dist:
        class ReduceOp:
            MAX = 0

        def barrier(self):
            pass

        def all_reduce(self, t, op=None):
            pass

    calls.clear()
    args = types.SimpleNamespace(m=m, k=k, L1=0.0, L2=0.0, solver="cholesky")
    r = bench.run_e2e_sharded(args, Eng(), Dist(), 3, 0, 2)
    assert r["value"] > 0 and r["unit"] == "nnz/s" and r["h2d_bytes_per_step"] > 0 and r["d2h_bytes_per_step"] > 0
    # block-wise factor I/O: only this rank's rows of W_T / H cross PCIe, both ways (warm-up call + timed call)
    assert calls == [("set_factor_blocks", (30, k), (20, k)), "get_factor_blocks"] * 2
    assert r["d2h_bytes_per_step"] == (30 * k * 4 + 20 * k * 4 + 4 * k) // 3


def test_row_block_operand_is_a_column_slice_of_the_transpose():
    """The premise of Engine::set_matrix_host_shard (in-process multi-GPU): the W half-step operand of rank g —
    (A[I_g, :])ᵀ as CSC with ascending global column ids, which set_matrix_sharded builds from a host-extracted row
    block — is the contiguous column range I_g of the full transpose with its pointers rebased; and the H half-step
    operand is the column range J_g of A itself."""
    import scipy.sparse as sp
    from rcppml_b200 import shard
    from helpers import random_csc
    m, n, world = 203, 117, 4
    A = random_csc(m, n, 0.07, 12, ragged=True)
    At = A.T.tocsc()
    At.sort_indices()
    for rank in range(world):
        r0, rc = shard.block_of(m, world, rank)
        rp, ri, rx = shard.extract_row_block(A.indptr, A.indices, A.data, r0, rc)       # A[I, :], local row ids
        ref = sp.csc_matrix((rx, ri, rp), shape=(rc, n)).T.tocsc()                       # what the engine transposes
        ref.sort_indices()
        lo, hi = At.indptr[r0], At.indptr[r0 + rc]
        assert np.array_equal(ref.indptr, At.indptr[r0:r0 + rc + 1] - lo)                # rebase_pointers_kernel
        assert np.array_equal(ref.indices, At.indices[lo:hi]) and np.array_equal(ref.data, At.data[lo:hi])
        c0, cc = shard.block_of(n, world, rank)
        cp, ci, cx = shard.extract_shard(A.indptr, A.indices, A.data, c0, cc)
        lo, hi = A.indptr[c0], A.indptr[c0 + cc]
        assert np.array_equal(cp, A.indptr[c0:c0 + cc + 1] - lo)
        assert np.array_equal(ci, A.indices[lo:hi]) and np.array_equal(cx, A.data[lo:hi])


def test_balanced_cuts_host_logic_and_library_agree():
    """Work-balanced contiguous partitions: rcppml_b200.shard.balanced_cuts — ascending cuts covering
    0 .. n, every block within one item's weight of the ideal share — and the library's own column cuts (the ones the
    in-process multi-GPU path takes from the caller's col_ptr) equal to it, on the reference's skewed datasets."""
    import ctypes as C
    from rcppml_b200 import _lib, shard
    lib = _lib.load()
    root = os.path.dirname(os.path.abspath(__file__))
    cases = []
    for name in ("pbmc3k_500x200.npz", "movielens.npz"):
        g = np.load(os.path.join(root, "golden", name))
        key = "indptr" if "indptr" in g.files else [f for f in g.files if f.endswith("p") or "ptr" in f][0]
        cases.append(np.asarray(g[key], dtype=np.int32))
    rng = np.random.default_rng(3)
    cases.append(np.concatenate([[0], np.cumsum(rng.integers(0, 2000, size=777))]).astype(np.int32))
    for indptr in cases:
        n = indptr.size - 1
        counts = np.diff(indptr)
        for world in (2, 3, 8):
            for per_item in (0, 16, 64):
                cuts = shard.balanced_cuts(counts, world, per_item=per_item)
                assert cuts[0] == 0 and cuts[-1] == n and np.all(np.diff(cuts) >= 0) and cuts.size == world + 1
                work = counts.astype(np.float64) + per_item
                ideal = work.sum() / world
                for r in range(world):
                    blk = work[cuts[r]:cuts[r + 1]].sum()
                    assert abs(blk - ideal) <= 2 * work.max() + 1e-9, (world, per_item, r, blk, ideal)
                out = np.zeros(world + 1, np.int32)
                assert lib.rcppml_b200_balanced_col_cuts(indptr.ctypes.data_as(C.POINTER(C.c_int)), n, world, per_item,
                                                         out.ctypes.data_as(C.POINTER(C.c_int))) == 0
                assert np.array_equal(out, cuts), (world, per_item, out, cuts)


def test_bench_parity_object_host_logic(oracle, monkeypatch):
    """bench.parity_vs_oracle (the N = 1 `parity` object of every bench line) with a stand-in engine: an engine that
    returns the oracle's own factors is reported ok with zero error and equal zero patterns; one whose W differs in a
    single entry beyond 1e-5, or whose zero pattern differs, is reported NOT ok. No CUDA."""
    import sys
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench
    import rcppml_b200 as rb
    m, n, k, iters = 400, 150, 8, 3
    Ap, Ai, Ax = oracle.synth_csc(m, n, 0, 0.05, bench.SEED_A)
    W0, H0 = oracle.initialize_factors(k, m, n, bench.SEED_INIT)
    ref = oracle.nmf_fit(Ap, Ai, Ax, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=1)
    cpu = {"_fit": ref, "_shape": (m, n), "_A": (Ap, Ai, Ax), "full_matrix": False, "sample": "leading block, test"}
    args = types.SimpleNamespace(k=k, L1=0.0, L2=0.0)

    def make_engine(W, H, d):
        class Eng:
            def __init__(self, dev):
                pass

            def set_matrix(self, *a):
                pass

            def set_factors(self, *a):
                pass

            def fit(self, cfg):
                return types.SimpleNamespace(status=0, iterations=cfg.max_iter)

            def get_factors(self):
                return W, H, d

            def loss_history(self, count):
                return ref.loss_history[:count]

            def cd_sweeps(self):
                return ref.cd_sweeps

            def close(self):
                pass
        return Eng

    monkeypatch.setattr(rb, "Engine", make_engine(ref.W_T.copy(), ref.H.copy(), ref.d.copy()))
    good = bench.parity_vs_oracle(args, None, cpu, 1)
    assert good["ok"] and good["max_rel_err"] == 0.0 and good["zero_pattern_equal"] and good["bit_identical_W"]
    assert good["rows_compared_W"] == m and good["rows_compared_H"] == n and good["iterations"] == iters
    W_bad = ref.W_T.copy()
    i, j = np.unravel_index(np.argmax(W_bad), W_bad.shape)
    W_bad[i, j] *= np.float32(1.001)
    monkeypatch.setattr(rb, "Engine", make_engine(W_bad, ref.H.copy(), ref.d.copy()))
    bad = bench.parity_vs_oracle(args, None, cpu, 1)
    assert not bad["ok"] and bad["rel_err"]["W"] > 1e-5 and not bad["bit_identical_W"]
    H_bad = ref.H.copy()
    zi = np.argwhere(H_bad == 0)
    if len(zi):
        H_bad[tuple(zi[0])] = np.float32(1e-12)            # inside the tolerance, but a different active set
        monkeypatch.setattr(rb, "Engine", make_engine(ref.W_T.copy(), H_bad, ref.d.copy()))
        pat = bench.parity_vs_oracle(args, None, cpu, 1)
        assert not pat["ok"] and not pat["zero_pattern_equal"] and pat["max_rel_err"] <= 1e-5
    # both arms name the workload identically, whatever the run-specific details
    a1 = bench.parse_args(["--gpus", "1"])
    a2 = bench.parse_args(["--impl", "reference", "--gpus", "8", "--steps", "20", "--warmup", "5"])
    assert bench.workload_string(a1) == bench.workload_string(a2)
    assert bench.parse_args(["--rows", "5", "--cols", "6", "--rank", "7"]).m == 5
