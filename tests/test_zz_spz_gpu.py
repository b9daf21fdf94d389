"""On-disk ingest on the GPU: the reference's own test file for this path re-expressed
(tests/testthat/test_gpu_sparsepress.R — st_read_gpu / st_free_gpu / zero-copy NMF from a .spz file), plus
file -> engine. The files are the golden set written by the reference's writer (tests/golden/spz) and, where
oracle/_ref travelled, the reference's real dataset. Runs last (file name): it is the newest path.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "spz")


def _device_arrays(g):
    """The three device arrays behind a gpu_sparse_matrix, copied back."""
    import ctypes as C
    import torch   # noqa: F401  (CUDA context; the copies below go through cudart)
    cudart = C.CDLL("libcudart.so.12", mode=C.RTLD_GLOBAL)
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    nnz = int(g.nnz)
    p = np.empty(g.n + 1, np.int32)
    i = np.empty(nnz, np.int32)
    x = np.empty(nnz, np.float64)
    for host, addr in ((p, g.col_ptr), (i, g.row_idx), (x, g.values)):
        if host.nbytes:
            assert cudart.cudaMemcpy(host.ctypes.data, int(addr), host.nbytes, 2) == 0      # cudaMemcpyDeviceToHost
    return p, i, x


@pytest.mark.parametrize("name", ["u8_t", "f32_t", "u16_escapes", "quant8_t", "f64", "f16", "u32_t"])
def test_st_read_gpu_leaves_the_matrix_on_the_device(name):
    """test_gpu_sparsepress.R:7-36 ("st_read_gpu reads .spz v2 to GPU memory") + the contents."""
    import rcppml_b200 as rb
    path = os.path.join(GOLDEN, name + ".spz")
    ref = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = rb.st_read_gpu(path)
    assert isinstance(g, rb.GpuSparseMatrix)
    assert (g.m, g.n, g.nnz, g.device) == (int(ref["m"]), int(ref["n"]), float(len(ref["r1_i"])), 0)
    assert g.col_ptr != 0 and g.row_idx != 0 and g.values != 0
    assert g.shape == (g.m, g.n) and "GPU Sparse Matrix" in str(g)               # :88-108 dim / print methods
    p, i, x = _device_arrays(g)
    assert np.array_equal(p, ref["r1_p"]) and np.array_equal(i, ref["r1_i"]) and np.array_equal(x, ref["r1_x"])
    assert rb.st_free_gpu(g) is None                                             # :40-60 "st_free_gpu frees GPU memory"
    assert g.col_ptr == 0 and g.row_idx == 0 and g.values == 0
    rb.st_free_gpu(g)                                                            # freeing twice is harmless


def test_st_read_gpu_and_st_free_gpu_errors(tmp_path):
    """test_gpu_sparsepress.R:63-77."""
    import rcppml_b200 as rb
    with pytest.raises(TypeError, match="gpu_sparse_matrix"):
        rb.st_free_gpu({"a": 1})
    with pytest.raises(TypeError, match="gpu_sparse_matrix"):
        rb.st_free_gpu("not_a_matrix")
    with pytest.raises(FileNotFoundError):
        rb.st_read_gpu("/nonexistent/path.spz")
    raw = open(os.path.join(GOLDEN, "u8_t.spz"), "rb").read()
    bad = tmp_path / "v3.spz"
    bad.write_bytes(raw[:4] + b"\x03\x00" + raw[6:])
    with pytest.raises(rb.streampress.SpzError) as ei:
        rb.st_read_gpu(str(bad))
    assert ei.value.status == 4                                                  # src/sp_gpu_bridge.cu:86-90
    bad.write_bytes(raw[:300])
    with pytest.raises(rb.streampress.SpzError) as ei:
        rb.st_read_gpu(str(bad))
    assert ei.value.status == 5


def test_nmf_from_a_gpu_sparse_matrix_matches_the_host_path():
    """test_gpu_sparsepress.R:111-190: zero-copy NMF on the decoded file — valid factors, and the same factors as the
    host-memory entry on the same matrix and initialisation (the R test allows 15 % on the MSE against the CPU fit; the
    two GPU entries here are the same engine and must agree exactly)."""
    import rcppml_b200 as rb
    path = os.path.join(GOLDEN, "u8_t.spz")
    ref = np.load(os.path.join(GOLDEN, "u8_t.npz"))
    m, n, k = int(ref["m"]), int(ref["n"]), 5
    g = rb.st_read_gpu(path)
    rng = np.random.default_rng(42)
    W0, H0 = rng.random((m, k)), rng.random((n, k))
    zc = rb.nmf_zerocopy(g, k, maxit=10, tol=1e-10, seed=42, w_init=W0)
    assert zc.status == 0 and zc.W_T.shape == (m, k) and zc.H.shape == (n, k) and zc.d.shape == (k,)
    assert np.all(np.isfinite(zc.W_T)) and np.all(np.isfinite(zc.H)) and np.all(zc.W_T >= 0) and np.all(zc.H >= 0)
    # same call with explicit H0, against the host entry
    zc2 = rb.gpu_nmf_zerocopy(g.col_ptr, g.row_idx, g.values, m, n, g.nnz, k, W0, H0, maxit=10, tol=0.0)
    host = rb.bridge_nmf_sparse(ref["r1_p"], ref["r1_i"], ref["r1_x"], m, n, k, W0.astype(np.float32), H0.astype(np.float32),
                                max_iter=10, tol=0.0, solver_mode=1)
    rb.st_free_gpu(g)
    assert zc2.status == 0 and host.status == 0
    assert np.array_equal(zc2.W_T, host.W_T) and np.array_equal(zc2.H, host.H) and np.array_equal(zc2.d, host.d)
    A = sp.csc_matrix((ref["r1_x"], ref["r1_i"], ref["r1_p"]), shape=(m, n)).toarray()
    mse = np.mean((A - (zc2.W_T * zc2.d) @ zc2.H.T) ** 2)
    assert mse < np.mean(A ** 2) * 0.95                                          # better than the zero model


@pytest.mark.parametrize("name", ["u8_t", "f32_t", "u16_escapes", "f32_rowsort_t", "quant8_t"])
def test_file_to_engine_equals_host_arrays_to_engine(name):
    """rcppml_b200_set_matrix_spz: the device operands (A and A^T) and a fit equal those of set_matrix on the decoded
    arrays. On one GPU the transpose section of the file is left alone by default (the device transpose is far cheaper
    than decoding it); when asked for it is used if usable (not under a row permutation; QUANT8 quantises the two
    sections per chunk — the transpose section is then still what the FILE says A^T is)."""
    import rcppml_b200 as rb
    path = os.path.join(GOLDEN, name + ".spz")
    ref = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, n, k = int(ref["m"]), int(ref["n"]), 6
    A = sp.csc_matrix((ref["r1_x"].astype(np.float32), ref["r1_i"], ref["r1_p"]), shape=(m, n))
    eng = rb.Engine(0)
    try:
        assert eng.set_matrix_spz(path) is False                                  # one GPU: the device transposes
        tp0, ti0, tx0 = eng.get_matrix_t()
        used = eng.set_matrix_spz(path, stored_transpose=True)
        assert used == (("t_p" in ref) and "rowsort" not in name)
        p, i, x = eng.get_matrix()
        assert np.array_equal(p, ref["r1_p"]) and np.array_equal(i, ref["r1_i"]) and np.array_equal(x, ref["r1_x"].astype(np.float32))
        tp, ti, tx = eng.get_matrix_t()
        if used:
            assert np.array_equal(tp, ref["t_p"]) and np.array_equal(ti, ref["t_i"]) and np.array_equal(tx, ref["t_x"].astype(np.float32))
        if name != "quant8_t" and "rowsort" not in name:
            At = A.T.tocsc()
            At.sort_indices()
            assert np.array_equal(tp, At.indptr) and np.array_equal(ti, At.indices) and np.array_equal(tx, At.data)
            assert np.array_equal(tp0, At.indptr) and np.array_equal(ti0, At.indices) and np.array_equal(tx0, At.data)
            eng.init_factors(k, 42)
            res = eng.fit(rb.make_config(k, max_iter=5, tol=0.0, solver_mode=1))
            from_file = eng.get_factors() + (eng.loss_history(5),)
            eng.set_matrix(m, n, A.indptr, A.indices, A.data)
            eng.init_factors(k, 42)
            res2 = eng.fit(rb.make_config(k, max_iter=5, tol=0.0, solver_mode=1))
            from_host = eng.get_factors() + (eng.loss_history(5),)
            assert res.status == 0 and res2.status == 0
            assert all(np.array_equal(a, b) for a, b in zip(from_file, from_host))
    finally:
        eng.close()


def test_the_reference_dataset_from_its_file_to_a_fit():
    """inst/extdata/pbmc3k.spz (13714 x 2700 counts, uint16 with escapes, no stored transpose) -> engine -> fit, against
    the same matrix decoded by the reference's reader (oracle/_ref/pbmc3k.bin)."""
    import rcppml_b200 as rb
    from helpers import load_pbmc3k
    path = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "pbmc3k.spz")
    A = load_pbmc3k()
    if A is None or not os.path.exists(path):
        pytest.skip("oracle/_ref/pbmc3k.{spz,bin} not built")
    m, n, k = A.shape[0], A.shape[1], 16
    eng = rb.Engine(0)
    try:
        assert eng.set_matrix_spz(path) is False
        p, i, x = eng.get_matrix()
        assert np.array_equal(p, A.indptr) and np.array_equal(i, A.indices) and np.array_equal(x, A.data)
        eng.init_factors(k, 42)
        eng.fit(rb.make_config(k, max_iter=4, tol=0.0, solver_mode=1))
        a = eng.get_factors()
        eng.set_matrix(m, n, A.indptr, A.indices, A.data)
        eng.init_factors(k, 42)
        eng.fit(rb.make_config(k, max_iter=4, tol=0.0, solver_mode=1))
        b = eng.get_factors()
        assert all(np.array_equal(u, v) for u, v in zip(a, b))
    finally:
        eng.close()
