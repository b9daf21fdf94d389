"""GPU parity tests: the CUDA engine (through the C ABI of RcppML_gpu.so) against the CPU oracle
on the same seeded inputs. Bit-exact for integer work (generator, transpose, RNG); 1e-5 relative
for fp32 factors (tests/helpers.py:RTOL), identical zero patterns (active sets)."""
import numpy as np
import pytest

from helpers import RTOL, random_csc, rel_err, zero_pattern_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import rcppml_b200 as rb
    e = rb.Engine(0)
    yield e
    e.close()


def test_detect_reports_b200():
    import rcppml_b200 as rb
    info = rb.gpu_detect()
    assert info["status"] == 0 and info["num_gpus"] >= 1
    assert info["total_mem_mb"][0] > 100_000


def test_fma_division_is_ieee_exact():
    """The solvers divide with two FMA corrections of a*RN(1/d); must equal __fdiv_rn bit for bit."""
    import ctypes as C
    from rcppml_b200 import _lib
    bad = C.c_int64(-1)
    _lib.check(_lib.load().rcppml_b200_selftest_division(2_000_000_000, 12345, C.byref(bad)), "selftest")
    assert bad.value == 0, bad.value


def test_synthetic_generator_bit_exact(eng):
    from rcppml_b200 import synth
    for (m, n_local, col_begin, dens) in [(5000, 300, 0, 0.01), (20000, 128, 77, 0.05), (100000, 64, 5, 0.03)]:
        eng.set_matrix_synthetic(m, n_local, col_begin, dens, synth.SEED_A)
        p, i, x = eng.get_matrix()
        hp, hi, hx = synth.synth_csc(m, n_local, col_begin, dens, synth.SEED_A)
        assert np.array_equal(p, hp) and np.array_equal(i, hi) and np.array_equal(x, hx)


def test_transpose_bit_exact(eng, oracle):
    A = random_csc(700, 400, 0.05, 3, ragged=True)
    eng.set_matrix(700, 400, A.indptr, A.indices, A.data)
    tp, ti, tx = eng.get_matrix_t()
    op, oi, ox = oracle.transpose_csc(A.indptr, A.indices, A.data, 700, 400)
    assert np.array_equal(tp, op) and np.array_equal(ti, oi) and np.array_equal(tx, ox)


def test_init_factors_bit_exact(eng, oracle):
    A = random_csc(300, 200, 0.05, 4)
    eng.set_matrix(300, 200, A.indptr, A.indices, A.data)
    for k, seed in [(7, 42), (64, 0), (20, 123456)]:
        eng.init_factors(k, seed)
        W, H, _ = eng.get_factors()
        oW, oH = oracle.initialize_factors(k, 300, 200, seed)
        assert np.array_equal(W, oW) and np.array_equal(H, oH)


@pytest.mark.parametrize("k", [6, 20, 32, 64, 128])
@pytest.mark.parametrize("solver", [0, 1])
def test_half_steps_match_oracle(eng, oracle, k, solver):
    import rcppml_b200 as rb
    m, n = 900, 500
    A = random_csc(m, n, 0.04, 10 + k, ragged=True)
    At = oracle.transpose_csc(A.indptr, A.indices, A.data, m, n)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    L1 = (0.01, 0.02)
    L2 = (0.03, 0.0)
    cfg = rb.make_config(k, L1=L1, L2=L2, solver_mode=solver, cd_maxit=100)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    for warm in (False, True):
        # ---- H update
        eng.set_factors(W0, H0)
        eng.half_step(cfg, 0, warm)
        _, H_gpu, d_gpu = eng.get_factors()
        G = oracle.gram(W0)
        G[np.diag_indices(k)] += np.float32(L2[1])
        H_ref = H0.copy()
        oracle.half_step(A.indptr, A.indices, A.data, W0, G, H_ref, solver_mode=solver, L1=L1[1], warm_start=warm)
        assert rel_err(H_gpu, H_ref) <= RTOL, (k, solver, warm, rel_err(H_gpu, H_ref))
        assert zero_pattern_equal(H_gpu, H_ref)
        d_ref = np.abs(H_ref.astype(np.float64)).sum(axis=0).astype(np.float32) + np.float32(1e-15)
        assert rel_err(d_gpu, d_ref) <= RTOL
        # ---- W update
        eng.set_factors(W0, H0)
        eng.half_step(cfg, 1, warm)
        W_gpu, _, _ = eng.get_factors()
        G = oracle.gram(H0)
        G[np.diag_indices(k)] += np.float32(L2[0])
        W_ref = W0.copy()
        oracle.half_step(At[0], At[1], At[2], H0, G, W_ref, solver_mode=solver, L1=L1[0], warm_start=warm)
        assert rel_err(W_gpu, W_ref) <= RTOL, (k, solver, warm, rel_err(W_gpu, W_ref))
        assert zero_pattern_equal(W_gpu, W_ref)


CASES = [
    # name, m, n, density, k, kwargs
    ("cd_k8", 500, 200, 0.08, 8, dict(solver_mode=0)),
    ("chol_k8", 500, 200, 0.08, 8, dict(solver_mode=1)),
    ("cd_k20_L1", 800, 610, 0.05, 20, dict(solver_mode=0, L1=(0.01, 0.01))),
    ("cd_k32", 1200, 700, 0.03, 32, dict(solver_mode=0)),
    ("chol_k64", 2000, 900, 0.03, 64, dict(solver_mode=1)),
    ("cd_k64_L1L2", 1500, 800, 0.03, 64, dict(solver_mode=0, L1=(0.01, 0.01), L2=(0.01, 0.01))),
    ("chol_k128_L1L2", 1500, 1100, 0.04, 128, dict(solver_mode=1, L1=(0.01, 0.01), L2=(0.01, 0.01))),
    ("chol_k32_ub_l2norm", 600, 400, 0.06, 32, dict(solver_mode=1, upper_bound=(0.02, 0.05), norm_type=1)),
    ("cd_k16_nonorm_seminmf", 600, 400, 0.06, 16, dict(solver_mode=0, norm_type=2, nonneg=(True, False))),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_full_fit_matches_oracle(eng, oracle, case):
    import rcppml_b200 as rb
    name, m, n, dens, k, kw = case
    iters = 8
    A = random_csc(m, n, dens, 7, counts=("k20" in name), ragged=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, threads=0, **kw)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    eng.set_factors(W0, H0)
    res = eng.fit(rb.make_config(k, max_iter=iters, tol=0.0, **kw))
    W, H, d = eng.get_factors()
    assert res.status == 0 and res.iterations == iters == ref.iterations
    errs = dict(W=rel_err(W, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d),
                loss=rel_err(eng.loss_history(iters), ref.loss_history))
    print(name, errs, "sweeps gpu/oracle", eng.cd_sweeps(), ref.cd_sweeps)
    assert errs["W"] <= RTOL and errs["H"] <= RTOL and errs["d"] <= RTOL, errs
    assert errs["loss"] <= 1e-5, errs
    assert zero_pattern_equal(W, ref.W_T) and zero_pattern_equal(H, ref.H)
    if kw.get("solver_mode", 0) == 0:
        assert eng.cd_sweeps() == ref.cd_sweeps


@pytest.mark.parametrize("k,solver", [(6, 0), (6, 1), (20, 0), (32, 1), (64, 0), (64, 1), (128, 1)])
def test_masked_fit_matches_oracle(eng, oracle, k, solver):
    """Explicit user mask (nmf/masked_nnls.hpp): per-column Gram correction + per-column solve + masked loss."""
    import rcppml_b200 as rb
    m, n, iters = 400, 260, 4
    A = random_csc(m, n, 0.08, 40 + k, ragged=True)
    M = random_csc(m, n, 0.04, 90 + k)
    M.sort_indices()
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    kw = dict(solver_mode=solver, L1=(0.01, 0.02), L2=(0.02, 0.01))
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0,
                         mask=(M.indptr, M.indices), **kw)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    eng.set_mask(M.indptr, M.indices)
    eng.set_factors(W0, H0)
    res = eng.fit(rb.make_config(k, max_iter=iters, tol=0.0, **kw))
    W, H, d = eng.get_factors()
    hist = eng.loss_history(iters)
    eng.set_mask(None)
    errs = dict(W=rel_err(W, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d), loss=rel_err(hist, ref.loss_history))
    print(k, solver, errs)
    assert res.iterations == iters and res.status == 0
    assert max(errs.values()) <= RTOL, errs
    assert zero_pattern_equal(W, ref.W_T) and zero_pattern_equal(H, ref.H)


def test_masked_abi_extension(oracle):
    import rcppml_b200 as rb
    m, n, k = 300, 200, 8
    A = random_csc(m, n, 0.1, 51)
    M = random_csc(m, n, 0.05, 52)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=5, tol=0.0, solver_mode=1,
                         mask=(M.indptr, M.indices))
    out = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=5, tol=0.0, solver_mode=1,
                               mask=(M.indptr, M.indices))
    assert out.status == 0
    assert rel_err(out.W_T, ref.W_T) <= RTOL and rel_err(out.H, ref.H) <= RTOL and rel_err(out.d, ref.d) <= RTOL
    assert abs(out.train_loss - ref.train_loss) <= 1e-5 * abs(ref.train_loss)


def test_pbmc3k_block_like_reference_gpu_test(eng, oracle):
    """tests/testthat/test_gpu_accuracy.R:222-251: pbmc3k[1:500,1:200], k=8, seed 42, maxit 100, tol 1e-10 — the
    reference accepts 10 % MSE / 0.95 cosine between its GPU and CPU; here the GPU must match the CPU restatement
    to 1e-5 with the same iteration count."""
    import rcppml_b200 as rb
    from helpers import load_pbmc3k_block
    A = load_pbmc3k_block()
    m, n, k = A.shape[0], A.shape[1], 8
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    for solver in (0, 1):
        ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=100, tol=1e-10, solver_mode=solver)
        eng.set_matrix(m, n, A.indptr, A.indices, A.data)
        eng.set_factors(W0, H0)
        res = eng.fit(rb.make_config(k, max_iter=100, tol=1e-10, solver_mode=solver))
        W, H, d = eng.get_factors()
        errs = dict(W=rel_err(W, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d))
        print("pbmc3k block solver", solver, res.iterations, ref.iterations, errs)
        assert res.iterations == ref.iterations and res.converged == ref.converged
        assert max(errs.values()) <= RTOL, errs
        # Stronger than the contract: the arithmetic definition is shared, so the factors are
        # bit-identical here. A fused multiply-add sneaking into the kernels (ptxas contracts packed f32x2
        # mul+add) would show up as a non-zero difference.
        assert np.array_equal(W, ref.W_T) and np.array_equal(H, ref.H) and np.array_equal(d, ref.d)


def test_pbmc3k_full_k32(eng, oracle):
    """BASELINE.json configs[2]: pbmc3k scRNA-seq, k=32 (mask="zeros" does not change a non-CV fit — SURVEY.md §8
    a12; R selects CD on GPU for k <= 32)."""
    import rcppml_b200 as rb
    from helpers import load_pbmc3k
    A = load_pbmc3k()
    if A is None:
        pytest.skip("oracle/_ref/pbmc3k.bin not built (needs /root/reference at build time)")
    m, n, k, iters = A.shape[0], A.shape[1], 32, 5
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    for solver in (0, 1):
        ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver)
        eng.set_matrix(m, n, A.indptr, A.indices, A.data)
        eng.set_factors(W0, H0)
        res = eng.fit(rb.make_config(k, max_iter=iters, tol=0.0, solver_mode=solver))
        W, H, d = eng.get_factors()
        errs = dict(W=rel_err(W, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d),
                    loss=rel_err(eng.loss_history(iters), ref.loss_history))
        print("pbmc3k k=32 solver", solver, errs, "loop ms/iter", res.loop_ms / iters)
        assert max(errs.values()) <= RTOL, errs
        assert zero_pattern_equal(W, ref.W_T) and zero_pattern_equal(H, ref.H)


def test_movielens_k20_l1(eng, oracle):
    """BASELINE.json configs[1]: the reference's movielens ratings matrix (data/movielens.rda, read without R by
    tests/golden/rdx3.py), k=20, L1=(0.01, 0.01); CD (what R selects for the GPU at k <= 32) and Cholesky."""
    import rcppml_b200 as rb
    from helpers import load_movielens
    A, frozen = load_movielens()
    m, n, k, iters = A.shape[0], A.shape[1], 20, 6
    assert (m, n, A.nnz) == (3867, 610, 75238)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    for solver in (0, 1):
        kw = dict(L1=(0.01, 0.01), solver_mode=solver)
        ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, **kw)
        eng.set_factors(W0, H0)
        res = eng.fit(rb.make_config(k, max_iter=iters, tol=0.0, **kw))
        W, H, d = eng.get_factors()
        errs = dict(W=rel_err(W, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d),
                    loss=rel_err(eng.loss_history(iters), ref.loss_history),
                    loss_frozen=rel_err(eng.loss_history(iters), frozen[f"loss_{solver}"]))
        print("movielens k=20 solver", solver, errs, "loop ms/iter", res.loop_ms / iters)
        assert max(errs.values()) <= RTOL, errs
        assert zero_pattern_equal(W, ref.W_T) and zero_pattern_equal(H, ref.H)
        if solver == 0:
            assert eng.cd_sweeps() == ref.cd_sweeps == int(frozen["sweeps_0"])


def test_aml_k6_through_the_sparse_path(eng, oracle):
    """BASELINE.json configs[0]: the reference quick-start matrix (data/aml.rda, dense 824 x 135), k=6, MSE. The
    reference runs it through its dense CPU path (out of scope here); this feeds the same numbers to the sparse
    kernels as CSC — every column is (almost) full, the opposite extreme of C4."""
    import rcppml_b200 as rb
    from helpers import load_aml_as_csc
    A = load_aml_as_csc()
    m, n, k, iters = 824, 135, 6, 10
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    for solver in (0, 1):
        ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver)
        eng.set_factors(W0, H0)
        eng.fit(rb.make_config(k, max_iter=iters, tol=0.0, solver_mode=solver))
        W, H, d = eng.get_factors()
        errs = dict(W=rel_err(W, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d),
                    loss=rel_err(eng.loss_history(iters), ref.loss_history))
        assert max(errs.values()) <= RTOL, (solver, errs)
        assert zero_pattern_equal(W, ref.W_T) and zero_pattern_equal(H, ref.H)


def test_convergence_and_patience(eng, oracle):
    import rcppml_b200 as rb
    m, n, k = 400, 300, 6
    A = random_csc(m, n, 0.1, 11)
    W0, H0 = oracle.initialize_factors(k, m, n, 1)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=200, tol=1e-4, solver_mode=1)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    eng.set_factors(W0, H0)
    res = eng.fit(rb.make_config(k, max_iter=200, tol=1e-4, solver_mode=1))
    assert ref.converged and res.converged
    assert res.iterations == ref.iterations
    assert abs(res.train_loss - ref.train_loss) <= 1e-5 * abs(ref.train_loss)


def test_reference_abi_unified_float(oracle):
    """Through the 73-pointer reference entry point, packed like bridge_nmf.hpp:199-342."""
    import rcppml_b200 as rb
    m, n, k = 700, 450, 20
    A = random_csc(m, n, 0.05, 21, counts=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    kw = dict(L1=(0.01, 0.01), solver_mode=0)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=6, tol=0.0, **kw)
    out = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=6, tol=0.0, **kw)
    assert out.status == 0 and out.iterations == 6
    assert rel_err(out.W_T, ref.W_T) <= RTOL and rel_err(out.H, ref.H) <= RTOL and rel_err(out.d, ref.d) <= RTOL
    assert abs(out.train_loss - ref.train_loss) <= 1e-5 * abs(ref.train_loss)
    # unsupported feature -> status -1, buffers untouched (the reference gateway then takes its CPU path)
    bad = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=2, L21=(0.1, 0.0))
    assert bad.status == -1 and np.array_equal(bad.W_T, W0)


def test_reference_abi_zerocopy_and_cached_engine(oracle):
    """rcppml_gpu_nmf_zerocopy_double (src/gpu_bridge_nmf.cu:879, R/gpu_backend.R:183): CSC arrays already on the
    device, addresses passed as doubles. Also exercises the process-cached engine behind the reference entry
    points: a big call, a small one, a masked one and a repeat must not leak state into each other."""
    import ctypes as C
    import torch
    import rcppml_b200 as rb
    from rcppml_b200 import _lib
    m, n, k = 900, 600, 32
    A = random_csc(m, n, 0.05, 77, counts=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=6, tol=0.0, solver_mode=1,
                         L1=(0.02, 0.01))
    first = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=6, tol=0.0, solver_mode=1,
                                 L1=(0.02, 0.01))
    assert first.status == 0 and rel_err(first.W_T, ref.W_T) <= RTOL and rel_err(first.H, ref.H) <= RTOL
    # smaller problem, other rank, masked, through the same cached engine
    m2, n2, k2 = 300, 200, 8
    A2 = random_csc(m2, n2, 0.1, 51)
    M2 = random_csc(m2, n2, 0.05, 52)
    W2, H2 = oracle.initialize_factors(k2, m2, n2, 42)
    ref2 = oracle.nmf_fit(A2.indptr, A2.indices, A2.data, m2, n2, k2, W2, H2, max_iter=4, tol=0.0, solver_mode=0,
                          mask=(M2.indptr, M2.indices))
    out2 = rb.bridge_nmf_sparse(A2.indptr, A2.indices, A2.data, m2, n2, k2, W2, H2, max_iter=4, tol=0.0, solver_mode=0,
                                mask=(M2.indptr, M2.indices))
    assert out2.status == 0 and rel_err(out2.W_T, ref2.W_T) <= RTOL and rel_err(out2.H, ref2.H) <= RTOL
    # zero-copy: device-resident CSC (the mask of the previous call must be gone)
    dp = torch.from_numpy(np.ascontiguousarray(A.indptr, np.int32)).cuda()
    di = torch.from_numpy(np.ascontiguousarray(A.indices, np.int32)).cuda()
    dx = torch.from_numpy(np.ascontiguousarray(A.data, np.float64)).cuda()
    torch.cuda.synchronize()
    zc = rb.gpu_nmf_zerocopy(dp.data_ptr(), di.data_ptr(), dx.data_ptr(), m, n, int(A.indptr[n]), k, W0, H0, maxit=6,
                             tol=0.0, L1=(0.01, 0.02))            # this wrapper's pairs are (H, W)
    assert zc.status == 0 and zc.iterations == 6
    assert np.array_equal(zc.W_T, first.W_T) and np.array_equal(zc.H, first.H) and np.array_equal(zc.d, first.d)
    assert zc.train_loss == first.train_loss
    ph = (C.c_double * 5)()
    lib = _lib.load()
    assert lib.rcppml_b200_last_call_phases(ph) == 0 and ph[3] > 0.0
    # host pointers are refused by the zero-copy entry; the cache survives a refused call and a release
    bad = rb.gpu_nmf_zerocopy(A.indptr.ctypes.data, A.indices.ctypes.data, 0, m, n, int(A.indptr[n]), k, W0, H0, maxit=2)
    assert bad.status == -1
    assert lib.rcppml_b200_release_cache() == 0
    again = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=6, tol=0.0, solver_mode=1,
                                 L1=(0.02, 0.01))
    assert np.array_equal(again.W_T, first.W_T) and np.array_equal(again.H, first.H)


@pytest.mark.parametrize("k,solver", [(6, 0), (20, 1), (64, 1), (64, 0), (128, 1)])
def test_row_panel_passes_are_bit_identical(eng, oracle, k, solver, monkeypatch):
    """Row-panel passes (engine.cu build_panels): the half-step split into P launches over row panels of the
    gathered factor must reproduce the single-launch fit bit for bit (same additions in the same CSC order),
    including empty segments, ragged columns and panels that split a column many times."""
    import rcppml_b200 as rb
    m, n = 1500, 700
    A = random_csc(m, n, 0.05, 300 + k, ragged=True)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    cfg = rb.make_config(k, max_iter=4, tol=0.0, solver_mode=solver, L1=(0.01, 0.0), L2=(0.0, 0.02))
    outs = []
    for panel_mb in ("0", "0.02", "0.004"):
        monkeypatch.setenv("RCPPML_B200_PANEL_MB", panel_mb)
        eng.init_factors(k, 42)
        res = eng.fit(cfg)
        assert res.status == 0 and res.iterations == 4
        outs.append(eng.get_factors() + (eng.loss_history(4),))
    monkeypatch.delenv("RCPPML_B200_PANEL_MB")
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=4, tol=0.0, solver_mode=solver,
                         L1=(0.01, 0.0), L2=(0.0, 0.02))
    assert rel_err(outs[1][0], ref.W_T) <= RTOL and rel_err(outs[1][1], ref.H) <= RTOL


@pytest.mark.parametrize("k", [5, 16, 20, 32, 50, 64, 100, 128])
@pytest.mark.parametrize("solver", [0, 1])
def test_tiled_kernel_is_bit_identical(eng, oracle, k, solver, monkeypatch):
    """kernels_tiled.cuh (wide gather -> shared-memory tile -> narrow solve -> whole-row stores) against the
    one-geometry kernels: factors, d and the loss history bit for bit (CD: sweep totals too), for ragged columns,
    L1/L2, an upper bound, both norms, short and long columns, and with row panels on top."""
    import rcppml_b200 as rb
    outs = {}
    for (m, n, dens) in ((1100, 600, 0.05), (300, 2100, 0.2)):            # W-update columns: ~30 / ~420 non-zeros
        A = random_csc(m, n, dens, 900 + k, ragged=True)
        eng.set_matrix(m, n, A.indptr, A.indices, A.data)
        for kw in (dict(L1=(0.01, 0.005), L2=(0.0, 0.01)), dict(upper_bound=(0.05, 0.08), norm_type=1)):
            cfg = rb.make_config(k, max_iter=4, tol=0.0, solver_mode=solver, cd_maxit=100, **kw)
            res = []
            for tiled, panel in (("0", None), ("2", None), ("2", "0.02")):
                monkeypatch.setenv("RCPPML_B200_TILED", tiled)
                if panel:
                    monkeypatch.setenv("RCPPML_B200_PANEL_MB", panel)
                eng.init_factors(k, 42)
                r = eng.fit(cfg)
                assert r.status == 0 and r.iterations == 4
                res.append(eng.get_factors() + (eng.loss_history(4), eng.cd_sweeps()))
                if panel:
                    monkeypatch.delenv("RCPPML_B200_PANEL_MB")
            monkeypatch.delenv("RCPPML_B200_TILED")
            for other in res[1:]:
                for a, b in zip(res[0][:4], other[:4]):
                    assert np.array_equal(a, b), (k, solver, m, kw)
                assert res[0][4] == other[4]
            outs[(m, tuple(kw))] = res[1]
    m, n = 1100, 600
    A = random_csc(m, n, 0.05, 900 + k, ragged=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=4, tol=0.0, solver_mode=solver,
                         L1=(0.01, 0.005), L2=(0.0, 0.01))
    got = outs[(1100, ("L1", "L2"))]
    assert rel_err(got[0], ref.W_T) <= RTOL and rel_err(got[1], ref.H) <= RTOL and rel_err(got[2], ref.d) <= RTOL


@pytest.mark.parametrize("k", [40, 64])
def test_tiled_cta768_and_hybrid_gather_are_bit_identical(eng, k, monkeypatch):
    """RCPPML_B200_TILED_CTA: the tiled Cholesky kernel as ONE 768-thread CTA per SM (L / Lt once per SM), and that layout
    with the hybrid gather (half of every segment's factor rows in registers, half through a cp.async shared-memory ring)
    must reproduce the default 256-thread kernel bit for bit — ragged short columns, both half-steps, L1/L2, a bound."""
    import rcppml_b200 as rb
    for (m, n, dens) in ((1500, 700, 0.04), (4000, 900, 0.03)):
        A = random_csc(m, n, dens, 1200 + k, ragged=True)
        eng.set_matrix(m, n, A.indptr, A.indices, A.data)
        for kw in (dict(L1=(0.01, 0.005), L2=(0.0, 0.01)), dict(upper_bound=(0.05, 0.08), norm_type=1)):
            cfg = rb.make_config(k, max_iter=4, tol=0.0, solver_mode=1, **kw)
            res = []
            monkeypatch.setenv("RCPPML_B200_TILED", "2")
            for mode in ("0", "1", "2"):
                monkeypatch.setenv("RCPPML_B200_TILED_CTA", mode)
                eng.init_factors(k, 42)
                r = eng.fit(cfg)
                assert r.status == 0 and r.iterations == 4
                res.append(eng.get_factors() + (eng.loss_history(4),))
            monkeypatch.delenv("RCPPML_B200_TILED_CTA")
            monkeypatch.delenv("RCPPML_B200_TILED")
            for other in res[1:]:
                for a, b in zip(res[0], other):
                    assert np.array_equal(a, b), (k, m, kw)


CD_GEOMS = {16: (301, 102, 4), 32: (701, 302, 104), 64: (702, 304, 108), 128: (704, 308, 116)}


@pytest.mark.parametrize("k", [5, 16, 20, 32, 50, 64, 100, 128])
def test_cd_kernel_geometries_are_bit_identical(eng, oracle, k, monkeypatch):
    """The coordinate-descent kernel (kernels_cd.cuh: narrow lane groups, pivots blocked by 4, branch-free
    steps, tolerance quotients split over lanes) must reproduce half_step_kernel<CD> bit for bit in every
    lane-group geometry — factors, d, loss history AND the total number of CD sweeps — and match the oracle."""
    import rcppml_b200 as rb
    m, n = 1100, 600
    A = random_csc(m, n, 0.05, 500 + k, ragged=True)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    kw = dict(solver_mode=0, L1=(0.01, 0.005), L2=(0.0, 0.01), cd_maxit=100)
    iters = 5
    cfg = rb.make_config(k, max_iter=iters, tol=0.0, **kw)
    kp = 16 if k <= 16 else 32 if k <= 32 else 64 if k <= 64 else 128
    outs = []
    monkeypatch.setenv("RCPPML_B200_TILED", "0")                # this test is about cd_half_step_kernel
    monkeypatch.setenv("RCPPML_B200_CD_KERNEL", "1")            # the original kernel
    eng.init_factors(k, 42)
    res = eng.fit(cfg)
    assert res.status == 0 and res.iterations == iters
    outs.append(eng.get_factors() + (eng.loss_history(iters), eng.cd_sweeps()))
    monkeypatch.delenv("RCPPML_B200_CD_KERNEL")
    for geom in CD_GEOMS[kp]:
        monkeypatch.setenv("RCPPML_B200_CD_GEOM", str(geom))
        eng.init_factors(k, 42)
        res = eng.fit(cfg)
        assert res.status == 0 and res.iterations == iters
        outs.append(eng.get_factors() + (eng.loss_history(iters), eng.cd_sweeps()))
    monkeypatch.delenv("RCPPML_B200_CD_GEOM")
    monkeypatch.delenv("RCPPML_B200_TILED")
    for gi, other in enumerate(outs[1:]):
        for a, b in zip(outs[0][:4], other[:4]):
            assert np.array_equal(a, b), (k, CD_GEOMS[kp][gi])
        assert outs[0][4] == other[4], (k, CD_GEOMS[kp][gi], outs[0][4], other[4])
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=0.0, **kw)
    assert rel_err(outs[1][0], ref.W_T) <= RTOL and rel_err(outs[1][1], ref.H) <= RTOL
    assert outs[1][4] == ref.cd_sweeps


@pytest.mark.parametrize("solver", [0, 1])
def test_cuda_graph_iterations_are_bit_identical(eng, oracle, solver, monkeypatch):
    """iterate() replays the steady-state iteration as a CUDA graph (engine.cu capture_iteration_graph). Same kernels,
    same order: factors, loss history, iteration count and the device-side convergence stop must be bit-identical
    to plain launches — with tol = 0 and with early stopping in the middle of the replayed iterations."""
    import rcppml_b200 as rb
    m, n, k = 900, 400, 20
    A = random_csc(m, n, 0.05, 77, ragged=True)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    for tol, iters in ((0.0, 12), (3e-3, 60)):
        cfg = rb.make_config(k, max_iter=iters, tol=tol, solver_mode=solver, L1=(0.01, 0.01), patience=2)
        outs = []
        for graph in ("0", "1"):
            monkeypatch.setenv("RCPPML_B200_GRAPH", graph)
            eng.init_factors(k, 42)
            res = eng.fit(cfg)
            assert res.status == 0
            outs.append(eng.get_factors() + (eng.loss_history(res.iterations), res.iterations, res.converged, res.gpu_launches))
        monkeypatch.delenv("RCPPML_B200_GRAPH")
        for a, b in zip(outs[0][:4], outs[1][:4]):
            assert np.array_equal(a, b)
        assert outs[0][4:] == outs[1][4:], (outs[0][4:], outs[1][4:])
        if tol > 0:
            assert outs[0][5] and outs[0][4] < iters        # stopped early, on the device
    # a second fit on new factors must not replay the previous fit's graph
    W0, H0 = oracle.initialize_factors(k, m, n, 7)
    eng.set_factors(W0, H0)
    cfg = rb.make_config(k, max_iter=6, tol=0.0, solver_mode=solver)
    res = eng.fit(cfg)
    ref = oracle.nmf_fit(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=6, tol=0.0, solver_mode=solver)
    W, H, d = eng.get_factors()
    assert res.iterations == 6 and rel_err(W, ref.W_T) <= RTOL and rel_err(H, ref.H) <= RTOL


def test_large_synthetic_properties(eng):
    """Full-width rows at reduced column count: size-independent properties (non-negativity,
    unit L1 row norms, monotone loss, determinism run to run)."""
    import rcppml_b200 as rb
    from rcppml_b200 import synth
    m, n, k = 200_000, 20_000, 64
    eng.set_matrix_synthetic(m, n, 0, 1e-3, synth.SEED_A)
    outs = []
    for _ in range(2):
        eng.init_factors(k, 42)
        res = eng.fit(rb.make_config(k, max_iter=5, tol=0.0, solver_mode=1))
        W, H, d = eng.get_factors()
        hist = eng.loss_history(5)
        outs.append((W, H, d, hist))
        assert res.status == 0 and W.min() >= 0 and H.min() >= 0
        assert np.allclose(W.sum(axis=0, dtype=np.float64), 1.0, atol=1e-4)
        assert np.all(np.diff(hist) <= 1e-6 * hist[0])
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("solver,density", [(1, 4e-3), (0, 1e-3)])
def test_c4_shaped_fit_matches_oracle_with_the_default_kernel_policy(eng, oracle, solver, density):
    """The configuration the headline is quoted on, at a quarter of its rows and columns and with NOTHING forced:
    250 K x 25 K, k = 64; Cholesky with C4's column statistics (1000 non-zeros per column, ~100 per row: 2.5e7 non-zeros —
    the H half-step on half_step_kernel<16,1>, the W half-step on the tiled kernel with 8-column batches, the dynamic
    work counters, CUDA-graph replay), coordinate descent on a quarter of that density (the CPU side of a CD fit is what
    bounds the test) — against the oracle on the same generator matrix: W, H, d, loss history to 1e-5 (in practice
    bit for bit), equal zero patterns, and in CD mode equal sweep totals. (bench.py makes the same comparison on the
    full 1e8-non-zero matrix.)"""
    import rcppml_b200 as rb
    from rcppml_b200 import synth
    m, n, k, iters = 250_000, 25_000, 64, 3
    eng.set_matrix_synthetic(m, n, 0, density, synth.SEED_A)
    Ap, Ai, Ax = eng.get_matrix()
    hp, hi, hx = oracle.synth_csc(m, n, 0, density, synth.SEED_A)
    assert np.array_equal(Ap, hp) and np.array_equal(Ai, hi) and np.array_equal(Ax, hx)
    eng.init_factors(k, 42)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    cfg = rb.make_config(k, max_iter=iters, tol=0.0, solver_mode=solver, cd_maxit=100)
    res = eng.fit(cfg)
    W, H, d = eng.get_factors()
    hist = eng.loss_history(iters)
    ref = oracle.nmf_fit(hp, hi, hx, m, n, k, W0, H0, max_iter=iters, tol=0.0, solver_mode=solver, cd_maxit=100)
    assert res.status == 0 and res.iterations == iters
    assert rel_err(W, ref.W_T) <= RTOL and rel_err(H, ref.H) <= RTOL and rel_err(d, ref.d) <= RTOL
    assert rel_err(hist, ref.loss_history) <= RTOL
    assert np.array_equal(W == 0, ref.W_T == 0) and np.array_equal(H == 0, ref.H == 0)
    if solver == 0:
        assert eng.cd_sweeps() == ref.cd_sweeps


def test_pre_stored_transpose_skips_the_device_transpose(eng):
    """rcppml_b200_set_matrix_with_transpose (a .spz file can carry CSC(A^T)): handing the engine the
    transpose it would have built gives the same device operands and the same fit, bit for bit."""
    import rcppml_b200 as rb
    m, n, k = 2100, 900, 20
    A = random_csc(m, n, 0.04, 5, ragged=True)
    At = A.T.tocsc()
    At.sort_indices()
    outs = []
    for pre in (False, True):
        if pre:
            eng.set_matrix_with_transpose(m, n, (A.indptr, A.indices, A.data), (At.indptr, At.indices, At.data))
        else:
            eng.set_matrix(m, n, A.indptr, A.indices, A.data)
        tp, ti, tx = eng.get_matrix_t()
        assert np.array_equal(tp, At.indptr) and np.array_equal(ti, At.indices) and np.array_equal(tx, At.data)
        eng.init_factors(k, 42)
        res = eng.fit(rb.make_config(k, max_iter=5, tol=0.0, solver_mode=1, L1=(0.01, 0.01)))
        assert res.status == 0
        outs.append(eng.get_factors() + (eng.loss_history(5),))
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))


def test_multi_gpu_matches_single_gpu():
    """Column-sharded fit over NCCL (tests/multigpu_check.py) — needs >= 2 GPUs on the box."""
    import os
    import subprocess
    import sys
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTIGPU_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("k", [5, 20, 64, 100])
def test_predict_nnls_match_reference_semantics(oracle, k):
    """predict()/nnls() on the GPU (rcppml_gpu_nnls_double) vs Rcpp_predict / c_nnls restated in fp64."""
    from rcppml_b200 import project
    m, n = 700, 300
    A = random_csc(m, n, 0.06, 70 + k, counts=True, ragged=True)
    rng = np.random.default_rng(k)
    w = rng.random((m, k))
    Ax64 = A.data.astype(np.float64)
    for kw in (dict(), dict(L1=0.05, L2=0.1), dict(upper_bound=0.3), dict(nonneg=False, L2=0.01)):
        ref = oracle.project_f64(A.indptr, A.indices, Ax64, m, n, w, **kw)            # (n, k)
        got = project.nnls(w, A, **kw)                                                 # (k, n)
        assert rel_err(got.T, ref) <= 1e-9, (k, kw, rel_err(got.T, ref))
        assert zero_pattern_equal(got.T, ref)
    ref = oracle.project_f64(A.indptr, A.indices, Ax64, m, n, w, L1=0.02)
    assert rel_err(project.predict(w, A, L1=0.02).T, ref) <= 1e-9
    # c_nnls warm start: B -= G h0, CD without tolerance (src/RcppFunctions_utils.cpp:346-356)
    h0 = rng.random((k, n)) * 0.1
    ref = oracle.project_f64(A.indptr, A.indices, Ax64, m, n, w, warm_start=h0.T, cd_maxit=7)
    got = project.nnls(w, A, warm_start=h0, cd_maxit=7)
    assert rel_err(got.T, ref) <= 1e-9


def test_evaluate_matches_dense_reconstruction(oracle):
    from rcppml_b200 import project
    m, n, k = 300, 200, 12
    A = random_csc(m, n, 0.1, 77, counts=True)
    rng = np.random.default_rng(5)
    w, h, d = rng.random((m, k)), rng.random((k, n)), rng.random(k) + 0.5
    for mz in (False, True):
        ref = oracle.evaluate_mse_f64(A.indptr, A.indices, A.data.astype(np.float64), m, n, w, d, h.T, mask_zeros=mz)
        got = project.evaluate(A, w, d, h, mask_zeros=mz)
        assert abs(got - ref) <= 1e-10 * abs(ref), (mz, got, ref)


CV_CASES = [
    # k, solver, mask_zeros, kwargs
    (6, 0, True, dict(cd_maxit=15)),
    (6, 1, True, dict()),
    (20, 1, True, dict(L1=(0.01, 0.02), L2=(0.02, 0.01))),
    (32, 0, True, dict(cd_maxit=10, L1=(0.01, 0.01))),
    (64, 1, True, dict()),
    (8, 1, False, dict()),
    (8, 0, False, dict(cd_maxit=10)),
    (128, 1, True, dict(L2=(0.01, 0.01))),
]


@pytest.mark.parametrize("k,solver,mask_zeros,kw", CV_CASES, ids=[f"k{c[0]}_s{c[1]}_mz{int(c[2])}" for c in CV_CASES])
def test_cv_fit_matches_oracle(eng, oracle, k, solver, mask_zeros, kw):
    """Speckled-mask cross-validation (nmf/fit_cv.hpp): hash mask bit-exact (same held-out count), factors and
    train/test loss histories within 1e-5, same early-stopping iteration."""
    import rcppml_b200 as rb
    m, n, iters = 350, 240, 12
    A = random_csc(m, n, 0.12, 60 + k, counts=True, ragged=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    cd_maxit = kw.pop("cd_maxit", 100)
    ref = oracle.nmf_fit_cv(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=iters, tol=1e-6, solver_mode=solver,
                            mask_zeros=mask_zeros, holdout_fraction=0.1, cv_seed=7, seed=42, cd_maxit=cd_maxit, **kw)
    eng.set_matrix(m, n, A.indptr, A.indices, A.data)
    eng.set_factors(W0, H0)
    res, cv = eng.fit_cv(rb.make_config(k, max_iter=iters, tol=1e-6, solver_mode=solver, cd_maxit=cd_maxit, **kw),
                         holdout_fraction=0.1, cv_seed=7, seed=42, mask_zeros=mask_zeros)
    W, H, d = eng.get_factors()
    assert res.status == 0
    assert cv["n_test"] == ref.n_test                       # integer hash mask: bit-exact
    assert res.iterations == ref.iterations and res.converged == ref.converged and cv["best_iter"] == ref.best_iter
    errs = dict(W=rel_err(W, ref.W_T), H=rel_err(H, ref.H), d=rel_err(d, ref.d),
                test=rel_err(cv["test_history"], ref.test_history), train=rel_err(cv["train_history"], ref.train_history))
    print(k, solver, mask_zeros, res.iterations, errs)
    assert max(errs["W"], errs["H"], errs["d"], errs["test"]) <= RTOL, errs
    assert errs["train"] <= 1e-4, errs                      # Gram-trick train loss: difference of large sums
    assert abs(cv["best_test_loss"] - ref.best_test_loss) <= 1e-5 * abs(ref.best_test_loss)


def test_cv_reference_abi(oracle):
    """Through the 51-pointer rcppml_gpu_nmf_cv_unified_float (gpu/bridge_nmf.hpp:78-99)."""
    import rcppml_b200 as rb
    m, n, k = 300, 220, 10
    A = random_csc(m, n, 0.1, 99, counts=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    ref = oracle.nmf_fit_cv(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=8, tol=1e-7, solver_mode=1,
                            mask_zeros=True, holdout_fraction=0.1, cv_seed=0, seed=42)
    out = rb.bridge_nmf_cv_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=8, tol=1e-7, solver_mode=1,
                                  mask_zeros=True, holdout_fraction=0.1, cv_seed=0, seed=42)
    assert out.status == 0 and out.iterations == ref.iterations and out.best_iter == ref.best_iter
    assert rel_err(out.W_T, ref.W_T) <= RTOL and rel_err(out.H, ref.H) <= RTOL and rel_err(out.d, ref.d) <= RTOL
    assert abs(out.test_loss - ref.test_loss) <= 1e-5 * abs(ref.test_loss)
    assert abs(out.best_test_loss - ref.best_test_loss) <= 1e-5 * abs(ref.best_test_loss)


def test_in_process_multi_gpu_behind_the_reference_entry_is_bit_identical():
    """RCPPML_NUM_GPUS (SURVEY.md §8b "Multi-GPU knob"): the single-process reference caller gets the sharded
    peer-memory loop with one engine + one host thread per device (abi_reference.cu). Needs >= 2 GPUs with peer
    access; run in a child process so that a failure cannot wedge the suite."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    code = r'''
import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import rcppml_b200 as rb
from rcppml_b200 import _lib
from helpers import random_csc
for (m, n, k, solver, kw) in [(900, 500, 16, 0, {}), (1201, 777, 64, 1, dict(L1=(0.01, 0.02), L2=(0.0, 0.01))), (640, 333, 8, 1, dict(upper_bound=(0.2, 0.3)))]:
    A = random_csc(m, n, 0.05, 5 + k, ragged=True)
    rng = np.random.default_rng(k)
    W0, H0 = rng.random((m, k)), rng.random((n, k))
    os.environ.pop("RCPPML_NUM_GPUS", None)
    one = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=7, tol=0.0, solver_mode=solver, **kw)
    # RCPPML_B200_MC: NVSwitch multicast replication ("2": from the solve kernel, "1": from the Gram kernel; where
    # supported) / unicast peer stores ("0"); the engines
    # behind the entry point are cached, so the cache is dropped when the mode changes
    for G, mc in (("2", "2"), ("all", "2"), ("2", "1"), ("all", "1"), ("2", "0"), ("all", "0")):
        os.environ["RCPPML_NUM_GPUS"] = G
        os.environ["RCPPML_B200_MC"] = mc
        _lib.load().rcppml_b200_release_cache()
        many = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=7, tol=0.0, solver_mode=solver, **kw)
        many2 = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=7, tol=0.0, solver_mode=solver, **kw)
        assert one.status == 0 and many.status == 0 and many2.status == 0, (G, mc, one.status, many.status, many2.status)
        assert np.array_equal(one.W_T, many.W_T) and np.array_equal(one.H, many.H) and np.array_equal(one.d, many.d), (G, mc, m, n, k)
        assert np.array_equal(many2.W_T, many.W_T) and np.array_equal(many2.H, many.H), "second call on the cached engines differs"
        # (tr(AtA) is summed per device, then over devices: the loss may differ in its last bit, the factors may not)
        assert one.iterations == many.iterations and abs(one.train_loss - many.train_loss) <= 1e-6 * abs(one.train_loss)
# the cross-validation entry (51 pointers) and the masked extension entry shard behind the same knob
import scipy.sparse as sp
m, n, k = 1500, 900, 16
A = random_csc(m, n, 0.05, 77, ragged=True)
rng = np.random.default_rng(5)
W0, H0 = rng.random((m, k)), rng.random((n, k))
M = sp.random(m, n, density=0.01, format="csc", random_state=rng, dtype=np.float32); M.sort_indices()
mask = (M.indptr.astype(np.int32), M.indices.astype(np.int32))
os.environ.pop("RCPPML_NUM_GPUS", None); os.environ.pop("RCPPML_B200_MC", None)
_lib.load().rcppml_b200_release_cache()
for solver in (0, 1):
    cv1 = rb.bridge_nmf_cv_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=6, tol=0.0, solver_mode=solver, cd_maxit=15, holdout_fraction=0.1, cv_seed=3)
    mk1 = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=5, tol=0.0, solver_mode=solver, mask=mask)
    for mc in ("2", "0"):
        os.environ["RCPPML_NUM_GPUS"] = "2"; os.environ["RCPPML_B200_MC"] = mc
        _lib.load().rcppml_b200_release_cache()
        cv2 = rb.bridge_nmf_cv_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=6, tol=0.0, solver_mode=solver, cd_maxit=15, holdout_fraction=0.1, cv_seed=3)
        mk2 = rb.bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W0, H0, max_iter=5, tol=0.0, solver_mode=solver, mask=mask)
        assert cv1.status == 0 and cv2.status == 0 and mk1.status == 0 and mk2.status == 0
        assert np.array_equal(cv1.W_T, cv2.W_T) and np.array_equal(cv1.H, cv2.H) and np.array_equal(cv1.d, cv2.d), ("cv", solver, mc)
        assert cv1.best_iter == cv2.best_iter and abs(cv1.test_loss - cv2.test_loss) <= 1e-6 * abs(cv1.test_loss)
        assert np.array_equal(mk1.W_T, mk2.W_T) and np.array_equal(mk1.H, mk2.H) and np.array_equal(mk1.d, mk2.d), ("masked", solver, mc)
    os.environ.pop("RCPPML_NUM_GPUS", None); os.environ.pop("RCPPML_B200_MC", None)
    _lib.load().rcppml_b200_release_cache()
print("INPROCESS_MULTIGPU_OK")
'''
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0 and "INPROCESS_MULTIGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_factor_blocks_on_one_gpu_are_the_whole_factors(oracle):
    """rcppml_b200_set/get_factor_blocks_f32 with a single rank: the blocks are the whole factors, so the fit and the
    round trip must equal set_factors / get_factors bit for bit. (The multi-rank case lives in tests/multigpu_check.py.)"""
    import os
    import rcppml_b200 as rb
    m, n, k = 500, 300, 12
    A = random_csc(m, n, 0.05, 17, ragged=True)
    W0, H0 = oracle.initialize_factors(k, m, n, 42)
    outs = []
    for blocks in (False, True):
        e = rb.Engine(0)
        try:
            e.set_matrix(m, n, A.indptr, A.indices, A.data)
            (e.set_factor_blocks if blocks else e.set_factors)(W0, H0)
            e.fit(rb.make_config(k, max_iter=4, tol=0.0, solver_mode=1))
            outs.append(e.get_factor_blocks() if blocks else e.get_factors())
        finally:
            e.close()
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
