"""On-disk ingest: rcppml_b200/csrc/spz_reader.cpp against the reference's own StreamPress v2 codec.

* tests/golden/spz/*.spz were written by the reference's writer (compress_v2) and the .npz next to each holds what the
  reference's readers (decompress_v2 with / without reorder, decompress_v2_transpose) return — made by
  tests/golden/make_spz_golden.py; these run anywhere.
* where oracle/_ref/spz_ref_tool exists (built from /root/reference by `make -C oracle ref`; it travels to the GPU box)
  more files are written and read back by the reference on the spot, and the reference's real dataset is compared.
Everything here is host code inside RcppML_gpu.so: no GPU needed. Bit-exact: indices, pointers and values (as float64
bit patterns, and as the float32 the engine computes in).
"""
import glob
import os
import zlib

import numpy as np
import pytest
import scipy.sparse as sp

from rcppml_b200 import streampress as S
from spz_helpers import REF_TOOL, read_bin, ref_tool, write_bin

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "spz")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.spz")))
have_tool = pytest.mark.skipif(not os.path.exists(REF_TOOL), reason="oracle/_ref/spz_ref_tool not built")


def bits(x):
    return np.ascontiguousarray(x, np.float64).view(np.uint64)


def same_csc(got, p, i, x):
    return np.array_equal(got[0], p) and np.array_equal(got[1], i) and np.array_equal(bits(got[2]), bits(x))


def test_the_golden_set_is_there():
    assert len(CASES) >= 9


@pytest.mark.parametrize("name", CASES)
def test_golden_file_decodes_like_the_reference_reader(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    with S.SpzFile(os.path.join(GOLDEN, name + ".spz")) as f:
        assert f.shape == (int(g["m"]), int(g["n"]))
        assert f.raw.nnz == len(g["a_i"])
        assert f.crc32() == f.raw.stored_crc32                        # the writer's CRC over everything before the footer
        with open(f.path, "rb") as fh:
            assert f.crc32() == zlib.crc32(fh.read()[:-16])
        for tag, reorder in (("r1", True), ("r0", False)):
            assert same_csc(f.read(0, None, reorder, 0, np.float64), g[f"{tag}_p"], g[f"{tag}_i"], g[f"{tag}_x"]), tag
            p32, i32, x32 = f.read(0, None, reorder, 0, np.float32)
            assert np.array_equal(p32, g[f"{tag}_p"]) and np.array_equal(i32, g[f"{tag}_i"])
            # the engine's fp32 values are the reference's doubles rounded once (src/RcppFunctions_nmf.cpp:4-5)
            assert np.array_equal(x32.view(np.uint32), g[f"{tag}_x"].astype(np.float32).view(np.uint32))
        assert np.array_equal(f.col_counts(0), np.diff(g["r1_p"]))
        if "t_p" in g:
            assert f.info()["has_transpose"]
            assert np.all(np.diff(g["t_p"]) >= 0)                     # fixture free of the reference's empty-chunk defect
            assert same_csc(f.read(1, None, False, 0, np.float64), g["t_p"], g["t_i"], g["t_x"])
            assert np.array_equal(f.col_counts(1), np.diff(g["t_p"]))
        else:
            assert not f.info()["has_transpose"]
            with pytest.raises(S.SpzError) as ei:
                f.read(1)
            assert ei.value.status == 6


@pytest.mark.parametrize("name", [c for c in CASES if "rowsort" not in c])
def test_golden_file_holds_the_matrix_that_was_written(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    with S.SpzFile(os.path.join(GOLDEN, name + ".spz")) as f:
        p, i, x = f.read(0, None, True, 0, np.float64)
        assert np.array_equal(p, g["a_p"]) and np.array_equal(i, g["a_i"])
        vt = f.info()["value_type"]
        if vt in ("uint8", "uint16", "uint32", "float64"):
            assert np.array_equal(x, g["a_x"])
        elif vt == "float32":
            assert np.array_equal(x, g["a_x"].astype(np.float32).astype(np.float64))
        elif vt == "float16":
            assert np.allclose(x, g["a_x"], rtol=2e-3, atol=1e-3)     # 11-bit significand
        else:                                                         # quant8: 255 levels per chunk
            assert np.max(np.abs(x - g["a_x"])) <= (g["a_x"].max() - g["a_x"].min()) / 255 * 0.51 + 1e-5
        if f.info()["has_transpose"]:
            At = sp.csc_matrix((x, i, p), shape=f.shape).T.tocsc()
            At.sort_indices()
            tp, ti, tx = f.read(1, None, False, 0, np.float64)
            assert np.array_equal(tp, At.indptr) and np.array_equal(ti, At.indices)
            if vt != "quant8":                                        # quant8 scales are per chunk: sections differ
                assert np.array_equal(tx, At.data)


def test_row_sorted_file_applies_the_stored_permutation_like_decompress_v2():
    """sparsepress_v2.hpp:1093-1104 maps every decoded row index through the stored record; rows of a column are not
    re-sorted. With reorder off the indices are the writer's sorted-space rows (ascending-nnz order of rows)."""
    g = np.load(os.path.join(GOLDEN, "u8_rowsort.npz"))
    with S.SpzFile(os.path.join(GOLDEN, "u8_rowsort.spz")) as f:
        assert f.info()["row_sorted"] and f.raw.row_permutation_len == f.raw.m
        perm = f.row_permutation()
        assert sorted(perm.tolist()) == list(range(f.raw.m))
        _, i0, _ = f.read(0, None, False)
        _, i1, _ = f.read(0, None, True)
        assert np.array_equal(i1, perm[i0].astype(np.int32))
        assert np.array_equal(i0, perm[g["a_i"]].astype(np.int32))    # the writer stored perm[row] (sparsepress_v2.hpp:503-508)


@pytest.mark.parametrize("name", ["u8_t", "f32_t", "u16_escapes", "quant8_t", "f64", "f16"])
def test_any_column_range_equals_the_slice_of_the_full_decode(name):
    rng = np.random.default_rng(5)
    with S.SpzFile(os.path.join(GOLDEN, name + ".spz")) as f:
        for section in ((0, 1) if f.info()["has_transpose"] else (0,)):
            P, I, X = f.read(section, None, False, 0, np.float64)
            nc = f.section_cols(section)
            cc = f.raw.chunk_cols
            ranges = [(0, 0), (0, 1), (nc - 1, nc), (nc, nc), (0, nc), (cc - 1, cc + 1), (cc, min(nc, 2 * cc)), (1, nc - 1)]
            ranges += [tuple(sorted(rng.integers(0, nc + 1, 2).tolist())) for _ in range(25)]
            for c0, c1 in ranges:
                c0, c1 = max(0, min(c0, nc)), max(0, min(c1, nc))
                if c0 > c1:
                    continue
                p, i, x = f.read(section, (c0, c1), False, 2, np.float64)
                assert f.range_nnz(section, c0, c1) == P[c1] - P[c0]
                assert np.array_equal(p, P[c0:c1 + 1] - P[c0]), (section, c0, c1)
                assert np.array_equal(i, I[P[c0]:P[c1]]) and np.array_equal(bits(x), bits(X[P[c0]:P[c1]])), (section, c0, c1)
        with pytest.raises(S.SpzError) as ei:
            f.read(0, (0, f.raw.n + 1))
        assert ei.value.status == 7


def test_thread_count_does_not_change_the_result():
    with S.SpzFile(os.path.join(GOLDEN, "f32_t.spz")) as f:
        ref = f.read(0, None, True, 1, np.float64)
        for t in (2, 3, 0, 64):
            assert same_csc(f.read(0, None, True, t, np.float64), *ref)


def test_python_mirror_of_the_r_api():
    path = os.path.join(GOLDEN, "f32_t.spz")
    g = np.load(os.path.join(GOLDEN, "f32_t.npz"))
    info = S.st_info(path)                                            # names of Rcpp_sp_metadata (sparsepress_bridge.cpp:370-391)
    assert (info["rows"], info["cols"], info["version"], info["value_type"]) == (200, 100, 2, "float32")
    assert info["nnz"] == len(g["a_i"]) and info["has_transpose"] and info["chunk_cols"] == 32 and info["num_chunks"] == 4
    assert info["file_bytes"] == os.path.getsize(path) and info["ratio"] == pytest.approx(info["raw_bytes"] / info["file_bytes"])
    A = S.st_read(path)
    assert A.shape == (200, 100) and A.dtype == np.float64 and A.nnz == len(g["a_i"])
    assert np.array_equal(A.indptr, g["r1_p"]) and np.array_equal(A.data, g["r1_x"])
    At = S.st_read_transpose(path)
    assert At.shape == (100, 200) and abs(At - A.T).max() == 0
    B = S.st_read(path, cols=(10, 50))                                # exact columns (the reference's cols= returns chunks)
    assert B.shape == (200, 40) and abs(B - A[:, 10:50]).max() == 0
    with pytest.raises(S.SpzError, match="pre-stored transpose"):
        S.st_read_transpose(os.path.join(GOLDEN, "f64.spz"))


def test_damaged_files_are_errors_with_the_reference_status_codes(tmp_path):
    with pytest.raises(S.SpzError) as ei:
        S.SpzFile(str(tmp_path / "missing.spz"))
    assert ei.value.status == 1                                       # sp_gpu_bridge.cu:60
    raw = open(os.path.join(GOLDEN, "u8_t.spz"), "rb").read()

    def status_of(data):
        q = tmp_path / "x.spz"
        q.write_bytes(data)
        try:
            with S.SpzFile(str(q)) as f:
                f.read(0)
                f.read(1)
        except S.SpzError as ex:
            return ex.status
        return 0

    assert status_of(raw) == 0
    assert status_of(raw[:5]) == 3                                    # sp_gpu_bridge.cu:78
    assert status_of(raw[:4] + b"\x03\x00" + raw[6:]) == 4            # v3 (dense) container: sp_gpu_bridge.cu:86
    assert status_of(raw[:4] + b"\x01\x00" + raw[6:]) == 4            # v1
    assert status_of(b"NOPE" + raw[4:]) == 5                          # bad magic: deserialize throws -> decode error
    assert status_of(raw[:100]) == 5
    for cut in (130, 200, len(raw) // 2, len(raw) - 17):
        assert status_of(raw[:cut]) == 5, cut                         # truncated anywhere: an error, never a crash
    # a damaged FOOTER alone is tolerated, as by the reference's readers (decompress_v2 never looks at it)
    assert status_of(raw[:-1]) in (0, 5)


@pytest.mark.parametrize("name", ["u8_t", "f32_t", "u16_escapes", "quant8_t"])
def test_random_corruption_never_crashes_and_never_yields_out_of_range_indices(name, tmp_path):
    """Every byte of the file is attacker-controlled input to an ingest path that ends in GPU gathers: whatever the
    damage, the reader returns an error or a well-formed matrix (pointers monotone, indices inside the dimension)."""
    raw = bytearray(open(os.path.join(GOLDEN, name + ".spz"), "rb").read())
    rng = np.random.default_rng(11)
    q = tmp_path / "fuzz.spz"
    outcomes = {"ok": 0, "error": 0}
    for trial in range(150):
        data = bytearray(raw)
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(data)))
            data[pos] = int(rng.integers(0, 256))
        q.write_bytes(bytes(data))
        try:
            with S.SpzFile(str(q)) as f:
                for section in ((0, 1) if f.raw.has_transpose else (0,)):
                    p, i, x = f.read(section, None, True, 2, np.float32)
                    rows = f.raw.m if section == 0 else f.raw.n
                    assert p[0] == 0 and np.all(np.diff(p) >= 0) and p[-1] == len(i)
                    assert len(i) == 0 or (i.min() >= 0 and i.max() < rows)
            outcomes["ok"] += 1
        except S.SpzError as ex:
            assert ex.status in (3, 4, 5, 6, 7)
            outcomes["error"] += 1
    assert outcomes["ok"] + outcomes["error"] == 150


# ---- with the reference codec at hand ---------------------------------------------------------------------------------

def _matrix(rng, m, n, density, kind):
    A = sp.random(m, n, density=density, format="csc", random_state=rng, dtype=np.float64)
    A.data = {"u8": lambda d: np.floor(d * 200) + 1, "u16": lambda d: np.floor(d * 60000) + 1,
              "u32": lambda d: np.floor(d * 4e9) + 1, "f": lambda d: (d - 0.5) * 100}[kind](A.data)
    A.sort_indices()
    return A


@have_tool
def test_files_written_by_the_reference_writer_read_like_the_reference_readers(tmp_path):
    rng = np.random.default_rng(7)
    a_bin, a_spz, d_bin = (str(tmp_path / n) for n in ("a.bin", "a.spz", "d.bin"))
    checked = 0
    for (m, n, density) in [(50, 30, 0.2), (1000, 77, 0.05), (3000, 300, 0.01), (20, 300, 0.4)]:
        for kind, precision in [("u8", "auto"), ("u16", "auto"), ("u32", "auto"), ("f", "auto"), ("f", "fp16"),
                                ("f", "quant8"), ("f", "fp64"), ("u8", "fp32")]:
            for row_sort, transpose, chunk_cols in [(0, 0, 2048), (0, 1, 16), (1, 0, 7), (1, 1, 64)]:
                write_bin(a_bin, _matrix(rng, m, n, density, kind))
                assert ref_tool("encode", a_bin, a_spz, precision, row_sort, transpose, chunk_cols)[0] == 0
                with S.SpzFile(a_spz) as f:
                    for reorder in (1, 0):
                        assert ref_tool("decode", a_spz, d_bin, reorder)[0] == 0
                        _, _, p, i, x = read_bin(d_bin)
                        assert same_csc(f.read(0, None, bool(reorder), 3, np.float64), p, i, x), (m, n, kind, precision, row_sort, chunk_cols)
                    if transpose:
                        assert ref_tool("decodet", a_spz, d_bin)[0] == 0
                        _, _, p, i, x = read_bin(d_bin)
                        got = f.read(1, None, False, 2, np.float64)
                        assert np.array_equal(got[1], i) and np.array_equal(bits(got[2]), bits(x))
                        if np.all(np.diff(p) >= 0):     # see test_empty_chunks_...: the reference's pointers can be damaged
                            assert np.array_equal(got[0], p)
                checked += 1
    assert checked == 128


@have_tool
def test_empty_chunks_decode_correctly_where_the_reference_reader_damages_the_pointers(tmp_path):
    """The writer emits the gap stream of a chunk WITHOUT non-zeros as bare column counts, with no u32 size prefix
    (sparsepress_v2.hpp:94, the early return); the reference's readers nevertheless take the first four bytes for the
    prefix when the stream has at least four (:991-1003, :1385-1397) and read four column counts beyond its end —
    the pointers of the last columns of such a chunk come out wrong (here: not even monotone). This reader knows an
    empty chunk from its descriptor and returns the matrix that was written."""
    A = sp.random(40, 64, density=0.2, format="csc", random_state=np.random.default_rng(1), dtype=np.float64)
    A = A.tolil(); A[:, 16:32] = 0; A = A.tocsc(); A.eliminate_zeros(); A.sort_indices()   # chunk 1 of 4 is empty
    a_bin, a_spz, d_bin = (str(tmp_path / n) for n in ("a.bin", "a.spz", "d.bin"))
    write_bin(a_bin, A)
    assert ref_tool("encode", a_bin, a_spz, "fp64", 0, 0, 16)[0] == 0
    assert ref_tool("decode", a_spz, d_bin, 1)[0] == 0
    _, _, rp, ri, rx = read_bin(d_bin)
    with S.SpzFile(a_spz) as f:
        p, i, x = f.read(0, None, True, 0, np.float64)
    assert np.array_equal(p, A.indptr) and np.array_equal(i, A.indices) and np.array_equal(x, A.data)
    assert np.array_equal(i, ri) and np.array_equal(x, rx)            # the entries agree ...
    assert not np.array_equal(rp, A.indptr)                           # ... the reference's pointers do not
    assert np.array_equal(np.delete(rp, np.arange(28, 32)), np.delete(A.indptr, np.arange(28, 32)))   # only those 4


@have_tool
def test_the_reference_dataset_pbmc3k(tmp_path):
    from helpers import load_pbmc3k
    path = os.path.join(os.path.dirname(REF_TOOL), "pbmc3k.spz")
    A = load_pbmc3k()
    if A is None or not os.path.exists(path):
        pytest.skip("oracle/_ref/pbmc3k.{spz,bin} not built")
    with S.SpzFile(path) as f:
        assert f.info()["value_type"] == "uint16" and f.shape == (13714, 2700) and f.raw.nnz == 2282976
        assert f.crc32() == f.raw.stored_crc32
        p, i, x = f.read(0, None, True, 0, np.float32)
    assert np.array_equal(p, A.indptr) and np.array_equal(i, A.indices) and np.array_equal(x, A.data)
    assert x.max() == 419.0                                           # values above 255 travel as escapes


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_st_read_gpu_without_a_gpu_fails_loudly():
    """The device half of the ingest has no host fallback: without a GPU rcppml_sp_read_gpu reports status 5
    (the reference's catch-all, src/sp_gpu_bridge.cu:117-120) and hands out no pointers."""
    with pytest.raises(S.SpzError) as ei:
        S.st_read_gpu(os.path.join(GOLDEN, "u8_t.spz"))
    assert ei.value.status == 5
    with pytest.raises(FileNotFoundError):
        S.st_read_gpu("/nonexistent/path.spz")


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_ingest_every_rank_decodes_exactly_its_two_blocks(world):
    """rcppml_b200_set_matrix_spz after comm_init: rank r decodes columns J_r of the main section and columns I_r of the
    transpose section. Host-side premise of that path: the first IS extract_shard(A, J_r), the second IS the transpose
    of extract_row_block(A, I_r) (global column ids ascending — what Engine::transpose_csc would have produced on the
    device), and the cuts of a work-balanced partition come from the chunk tables without decoding anything."""
    from rcppml_b200 import shard
    k = 8
    with S.SpzFile(os.path.join(GOLDEN, "f32_t.spz")) as f:
        m, n = f.shape
        P, I, X = f.read(0)
        col_cuts = shard.balanced_cuts(f.col_counts(0), world, per_item=k)
        row_cuts = shard.balanced_cuts(f.col_counts(1), world, per_item=k)
        assert np.array_equal(col_cuts, shard.balanced_cuts(np.diff(P), world, per_item=k))
        rows_nnz = np.bincount(I, minlength=m)
        assert np.array_equal(row_cuts, shard.balanced_cuts(rows_nnz, world, per_item=k))
        seen_cols = seen_rows = 0
        for r in range(world):
            for cuts, equal in ((None, True), ((col_cuts, row_cuts), False)):
                if equal:
                    (cb, nl), (rb, ml) = shard.block_of(n, world, r), shard.block_of(m, world, r)
                else:
                    cb, nl, rb, ml = cuts[0][r], cuts[0][r + 1] - cuts[0][r], cuts[1][r], cuts[1][r + 1] - cuts[1][r]
                p, i, x = f.read(0, (cb, cb + nl))
                ep, ei, ex = shard.extract_shard(P, I, X, cb, nl)
                assert np.array_equal(p, ep) and np.array_equal(i, ei) and np.array_equal(x, ex)
                tp, ti, tx = f.read(1, (rb, rb + ml), reorder=False)
                bp, bi, bx = shard.extract_row_block(P, I, X, rb, ml)          # n columns, block-relative rows
                Bt = sp.csc_matrix((bx, bi, bp), shape=(ml, n)).T.tocsc()      # m_loc columns, global column ids
                Bt.sort_indices()
                assert np.array_equal(tp, Bt.indptr) and np.array_equal(ti, Bt.indices) and np.array_equal(tx, Bt.data)
                if not equal:
                    seen_cols += nl
                    seen_rows += ml
        assert (seen_cols, seen_rows) == (n, m)


@pytest.mark.parametrize("name", ["u16_escapes", "f64", "u8_rowsort"])
def test_row_block_of_a_file_without_transpose_section(name):
    """Sharded ingest, fallback: a file without a (usable) transpose section gives a rank its row block by a full decode
    filtered on the host — the operand contract of set_matrix_sharded (shard.extract_row_block)."""
    from rcppml_b200 import shard
    with S.SpzFile(os.path.join(GOLDEN, name + ".spz")) as f:
        P, I, X = f.read(0)
        m = f.raw.m
        for world in (1, 2, 5):
            total = 0
            for r in range(world):
                rb, ml = shard.block_of(m, world, r)
                p, i, x = f.row_block(rb, ml, threads=2)
                ep, ei, ex = shard.extract_row_block(P, I, X, rb, ml)
                assert np.array_equal(p, ep) and np.array_equal(i, ei) and np.array_equal(x, ex)
                total += len(i)
            assert total == len(I)
        with pytest.raises(S.SpzError):
            f.row_block(m - 1, 2)


@have_tool
def test_files_with_obs_and_var_tables_read_the_same(tmp_path):
    """The writer places the serialized obs / var tables between the transpose section and the metadata and records
    their offsets in the header's reserved bytes (sparsepress_v2.hpp:810-818, header_v2.hpp:150-163). The ingest does not
    read the tables; it must not trip over them either."""
    A = _matrix(np.random.default_rng(9), 300, 120, 0.05, "f")
    a_bin, plain, tabled, d_bin = (str(tmp_path / n) for n in ("a.bin", "p.spz", "t.spz", "d.bin"))
    write_bin(a_bin, A)
    assert ref_tool("encode", a_bin, plain, "auto", 1, 1, 32)[0] == 0
    assert ref_tool("encode", a_bin, tabled, "auto", 1, 1, 32, 777, 1234)[0] == 0
    assert os.path.getsize(tabled) == os.path.getsize(plain) + 777 + 1234
    with S.SpzFile(plain) as f0, S.SpzFile(tabled) as f1:
        assert not f0.info()["has_obs"] and not f0.info()["has_var"]
        assert f1.info()["has_obs"] and f1.info()["has_var"] and f1.crc32() == f1.raw.stored_crc32
        for section, reorder in ((0, True), (0, False), (1, False)):
            assert same_csc(f1.read(section, None, reorder, 0, np.float64), *f0.read(section, None, reorder, 0, np.float64))
        assert np.array_equal(f1.row_permutation(), f0.row_permutation()) and len(f1.row_permutation()) == 300
    assert ref_tool("decode", tabled, d_bin, 1)[0] == 0
    _, _, p, i, x = read_bin(d_bin)
    with S.SpzFile(tabled) as f1:
        assert same_csc(f1.read(0, None, True, 0, np.float64), p, i, x)
