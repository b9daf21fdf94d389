"""Host mirror of the reference's StreamPress reading API for the on-disk ingest row:
`st_info`, `st_read`, `st_read_transpose` (R/streampress.R:163-211 over src/sparsepress_bridge.cpp:226-395) and
`st_read_gpu` / `st_free_gpu` / `.gpu_nmf_zerocopy` (R/sp_gpu.R:53-141, R/gpu_backend.R:183-265) — all through the C ABI
of rcppml_b200/lib/RcppML_gpu.so (include/rcppml_gpu.h, part 3). The decoder is csrc/spz_reader.cpp; nothing here
decodes in Python and there is no fallback: a missing library raises.

Only the v2 sparse container is read (like the reference's GPU reader, src/sp_gpu_bridge.cu:84-90); v1 and the dense v3
format belong to parts of the package outside this path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib

VALUE_TYPE_NAMES = ("uint8", "uint16", "uint32", "float32", "float16", "quant8", "float64")   # header_v2.hpp:55-66
VALUE_TYPE_BYTES = (1, 2, 4, 4, 2, 1, 8)                                                      # header_v2.hpp:68-79
STATUS_TEXT = {1: "cannot open the file", 2: "cannot read the file", 3: "file too small", 4: "not a v2 .spz file",
               5: "decode error", 6: "the file has no pre-stored transpose", 7: "bad argument"}


class SpzInfo(C.Structure):
    """rcppml_b200_spz_info (include/rcppml_gpu.h)."""
    _fields_ = [
        ("m", C.c_int), ("n", C.c_int), ("nnz", C.c_int64),
        ("chunk_cols", C.c_int), ("num_chunks", C.c_int),
        ("value_type", C.c_int), ("row_sorted", C.c_int),
        ("has_transpose", C.c_int), ("transpose_chunks", C.c_int), ("transp_chunk_cols", C.c_int),
        ("has_obs", C.c_int), ("has_var", C.c_int), ("has_metadata", C.c_int),
        ("row_permutation_len", C.c_int),
        ("density", C.c_float),
        ("file_bytes", C.c_int64), ("transpose_offset", C.c_int64), ("metadata_offset", C.c_int64),
        ("metadata_bytes", C.c_int64),
        ("stored_crc32", C.c_uint32),
    ]


class SpzError(_lib.NativeLibraryError):
    def __init__(self, status: int, what: str):
        super().__init__(what)
        self.status = status


def _bind(lib):
    if getattr(lib, "_spz_bound", False):
        return lib
    H, ip = C.c_void_p, C.POINTER(C.c_int)
    lib.rcppml_b200_spz_open.argtypes = [C.c_char_p, C.POINTER(H)]
    lib.rcppml_b200_spz_close.argtypes = [H]
    lib.rcppml_b200_spz_close.restype = None
    lib.rcppml_b200_spz_get_info.argtypes = [H, C.POINTER(SpzInfo)]
    lib.rcppml_b200_spz_crc32.argtypes = [H, C.POINTER(C.c_uint32)]
    lib.rcppml_b200_spz_range_nnz.argtypes = [H, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    lib.rcppml_b200_spz_col_counts.argtypes = [H, C.c_int, C.c_int, ip]
    lib.rcppml_b200_spz_read_f32.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, C.POINTER(C.c_float)]
    lib.rcppml_b200_spz_read_f64.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, C.POINTER(C.c_double)]
    lib.rcppml_b200_spz_row_block_f32.argtypes = [H, C.c_int, C.c_int, C.c_int, ip, ip, C.POINTER(C.c_float), C.c_int64,
                                                  C.POINTER(C.c_int64)]
    lib.rcppml_b200_spz_metadata.argtypes = [H, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.rcppml_b200_set_matrix_spz.argtypes = [C.c_void_p, H, C.c_int, C.c_int, ip]
    dp = C.POINTER(C.c_double)
    for name in ("rcppml_sp_read_gpu", "rcppml_st_read_gpu"):
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(C.c_char_p), ip, dp, dp, dp, ip, ip, dp, ip]
        fn.restype = None
    for name in ("rcppml_sp_free_gpu", "rcppml_st_free_gpu"):
        fn = getattr(lib, name)
        fn.argtypes = [dp, dp, dp, ip]
        fn.restype = None
    lib._spz_bound = True
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = _lib.load().rcppml_b200_last_error()
        text = msg.decode() if msg else STATUS_TEXT.get(rc, "unknown error")
        raise SpzError(rc, f"{what} failed (status {rc}): {text}")


class SpzFile:
    """An open `.spz` v2 file (mapped by the library). Use as a context manager or call close()."""

    def row_block(self, row_begin: int, m_loc: int, threads: int = 0):
        return _row_block(self, row_begin, m_loc, threads)

    def __init__(self, path: str):
        self._lib = _bind(_lib.load())
        self._h = C.c_void_p()
        self.path = os.path.abspath(os.path.expanduser(path))
        _check(self._lib.rcppml_b200_spz_open(self.path.encode(), C.byref(self._h)), f"st_read: open {path}")
        info = SpzInfo()
        _check(self._lib.rcppml_b200_spz_get_info(self._h, C.byref(info)), "st_info")
        self.raw = info

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rcppml_b200_spz_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        self.close()

    # -- metadata ---------------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return (self.raw.m, self.raw.n)

    def info(self) -> dict:
        """The list Rcpp_sp_metadata returns for a v2 file (src/sparsepress_bridge.cpp:357-391), same names."""
        r = self.raw
        raw_bytes = 12 + (r.n + 1) * 4 + r.nnz * 4 + r.nnz * VALUE_TYPE_BYTES[r.value_type]
        return {
            "rows": r.m, "cols": r.n, "nnz": float(r.nnz), "density_pct": float(r.density) * 100.0,
            "file_bytes": float(r.file_bytes), "raw_bytes": float(raw_bytes),
            "ratio": raw_bytes / r.file_bytes if r.file_bytes > 0 else 0.0,
            "version": 2, "value_type": VALUE_TYPE_NAMES[r.value_type], "chunk_cols": r.chunk_cols,
            "num_chunks": r.num_chunks, "row_sorted": bool(r.row_sorted), "has_transpose": bool(r.has_transpose),
            "has_metadata": bool(r.has_metadata), "has_obs": bool(r.has_obs), "has_var": bool(r.has_var),
            "transp_chunk_cols": r.transp_chunk_cols, "transpose_offset": float(r.transpose_offset),
        }

    def crc32(self) -> int:
        out = C.c_uint32(0)
        _check(self._lib.rcppml_b200_spz_crc32(self._h, C.byref(out)), "spz crc32")
        return out.value

    def metadata(self, key: int) -> bytes:
        n = C.c_int64(0)
        _check(self._lib.rcppml_b200_spz_metadata(self._h, key, None, 0, C.byref(n)), "spz metadata")
        if n.value == 0:
            return b""
        buf = (C.c_ubyte * n.value)()
        _check(self._lib.rcppml_b200_spz_metadata(self._h, key, buf, n.value, C.byref(n)), "spz metadata")
        return bytes(buf)

    def names(self, which: str):
        """Row or column names stored in the metadata section (NUL-separated, header_v2.hpp:286-314); [] if none."""
        rec = self.metadata({"rows": 0, "cols": 1}[which])
        return [s.decode() for s in rec.split(b"\0")[:-1]] if rec else []

    def row_permutation(self) -> np.ndarray:
        return np.frombuffer(self.metadata(2), dtype=np.uint32).copy()

    # -- decoding ---------------------------------------------------------------------------------------------------
    def section_cols(self, section: int) -> int:
        return self.raw.n if section == 0 else self.raw.m

    def range_nnz(self, section: int, c0: int, c1: int) -> int:
        out = C.c_int64(0)
        _check(self._lib.rcppml_b200_spz_range_nnz(self._h, section, c0, c1, C.byref(out)), "spz range_nnz")
        return out.value

    def col_counts(self, section: int = 0, threads: int = 0) -> np.ndarray:
        """Non-zeros per column of A (section 0) or per row of A (section 1: columns of the stored transpose), read
        from the chunk tables without entropy decoding — the input of a work-balanced partition (shard.balanced_cuts)."""
        out = np.zeros(self.section_cols(section), dtype=np.int32)
        _check(self._lib.rcppml_b200_spz_col_counts(self._h, section, threads, out.ctypes.data_as(C.POINTER(C.c_int))),
               "spz col_counts")
        return out

    def read(self, section: int = 0, cols=None, reorder: bool = True, threads: int = 0, dtype=np.float32):
        """(indptr, indices, data) of columns [c0, c1) of A (section 0) or of the stored transpose (section 1);
        indptr is rebased to 0. dtype float32 (the engine's) or float64 (the R boundary's)."""
        c0, c1 = (0, self.section_cols(section)) if cols is None else (int(cols[0]), int(cols[1]))
        nnz = self.range_nnz(section, c0, c1)
        p = np.empty(c1 - c0 + 1, dtype=np.int32)          # every entry is written by the decode (and only once)
        i = np.empty(max(nnz, 1), dtype=np.int32)
        ip = C.POINTER(C.c_int)
        if np.dtype(dtype) == np.float64:
            x = np.empty(max(nnz, 1), dtype=np.float64)
            rc = self._lib.rcppml_b200_spz_read_f64(self._h, section, c0, c1, int(reorder), threads, p.ctypes.data_as(ip),
                                                    i.ctypes.data_as(ip), x.ctypes.data_as(C.POINTER(C.c_double)))
        else:
            x = np.empty(max(nnz, 1), dtype=np.float32)
            rc = self._lib.rcppml_b200_spz_read_f32(self._h, section, c0, c1, int(reorder), threads, p.ctypes.data_as(ip),
                                                    i.ctypes.data_as(ip), x.ctypes.data_as(C.POINTER(C.c_float)))
        _check(rc, "st_read")
        return p, i[:nnz], x[:nnz]


def _row_block(f: "SpzFile", row_begin: int, m_loc: int, threads: int = 0):
    """(indptr, indices, data) of A[row_begin : row_begin + m_loc, :] with block-relative rows, decoded from the MAIN
    section and filtered on the host — what a sharded ingest falls back to for a file without a transpose section."""
    cap = max(int(f.raw.nnz), 1)
    p = np.zeros(f.raw.n + 1, np.int32)
    i = np.zeros(cap, np.int32)
    x = np.zeros(cap, np.float32)
    nnz = C.c_int64(0)
    ip = C.POINTER(C.c_int)
    _check(f._lib.rcppml_b200_spz_row_block_f32(f._h, row_begin, m_loc, threads, p.ctypes.data_as(ip), i.ctypes.data_as(ip),
                                                x.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(nnz)), "spz row block")
    return p, i[:nnz.value].copy(), x[:nnz.value].copy()


def st_info(path: str) -> dict:
    """`st_info(path)` (R/streampress.R:208-211)."""
    with SpzFile(path) as f:
        return f.info()


def _as_csc(p, i, x, shape):
    import scipy.sparse as sp
    A = sp.csc_matrix((x, i, p), shape=shape)
    return A


def st_read(path: str, cols=None, reorder: bool = True, threads=None):
    """`st_read(path, cols, reorder, threads)` (R/streampress.R:163-170): the matrix as scipy CSC with float64 values
    (a dgCMatrix). `cols=(c0, c1)` is a 0-based half-open column range and returns exactly those columns — the
    reference's `cols=` is documented as non-functional (tests/testthat/test_spz_roundtrip_comprehensive.R:93-99: it
    returns whole chunks). threads=None mirrors the R default: one thread below 50 MB, every core above. Row and
    column names, when the file stores them, are attached as `.rownames` / `.colnames`."""
    with SpzFile(path) as f:
        if threads is None:
            threads = 1 if f.raw.file_bytes < 50e6 else 0
        p, i, x = f.read(0, cols, reorder, int(threads), np.float64)
        ncols = f.raw.n if cols is None else int(cols[1]) - int(cols[0])
        A = _as_csc(p, i, x, (f.raw.m, ncols))
        A.rownames, A.colnames = f.names("rows"), f.names("cols")
        return A


def st_read_transpose(path: str, threads: int = 0):
    """`st_read_transpose(path)` (R/streampress.R:179-182; src/sparsepress_bridge.cpp:273-286): the pre-stored CSC(Aᵀ),
    n x m. Raises SpzError (status 6) when the file carries none."""
    with SpzFile(path) as f:
        if not f.raw.has_transpose:
            raise SpzError(6, "File does not contain a pre-stored transpose. "
                              "Use sp_write(..., include_transpose = TRUE) to create one.")
        p, i, x = f.read(1, None, False, threads, np.float64)
        return _as_csc(p, i, x, (f.raw.n, f.raw.m))


@dataclass
class GpuSparseMatrix:
    """The `gpu_sparse_matrix` object of R/sp_gpu.R:71-83: dimensions plus three device addresses (int32 col_ptr,
    int32 row_idx, float64 values), zeroed by st_free_gpu."""
    m: int
    n: int
    nnz: float
    device: int
    col_ptr: float
    row_idx: float
    values: float
    _freed: bool = field(default=False, repr=False)

    @property
    def shape(self):
        return (self.m, self.n)

    def __str__(self):
        # print.gpu_sparse_matrix (R/sp_gpu.R:166-173)
        return (f"GPU Sparse Matrix ({self.m} x {self.n}, nnz = {self.nnz:.0f})\n  Device: GPU {self.device}\n"
                f"  Memory: ~{(self.nnz * 8 + (self.n + 1) * 4) / 1024 ** 2:.1f} MB on device")

    def __del__(self):
        try:
            if self.col_ptr:
                st_free_gpu(self)
        except Exception:
            pass


def st_read_gpu(path: str, device: int = 0) -> GpuSparseMatrix:
    """`st_read_gpu(path, device)` (R/sp_gpu.R:53-109): decode a v2 file and leave it on the device."""
    lib = _bind(_lib.load())
    full = os.path.abspath(os.path.expanduser(path))
    if not os.path.exists(full):
        raise FileNotFoundError(path)                       # normalizePath(mustWork = TRUE)
    pth = C.c_char_p(full.encode())
    dev, m, n, st = C.c_int(int(device)), C.c_int(0), C.c_int(0), C.c_int(0)
    a, b, c, nnz = C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0)
    lib.rcppml_sp_read_gpu(C.byref(pth), C.byref(dev), C.byref(a), C.byref(b), C.byref(c), C.byref(m), C.byref(n),
                           C.byref(nnz), C.byref(st))
    if st.value != 0:
        raise SpzError(st.value, f"GPU decode failed with status {st.value}. Ensure file is .spz v2 format.")
    return GpuSparseMatrix(m.value, n.value, nnz.value, int(device), a.value, b.value, c.value)


def st_free_gpu(x: GpuSparseMatrix) -> None:
    """`st_free_gpu(x)` (R/sp_gpu.R:126-152): frees the three device arrays and zeroes the addresses."""
    if not isinstance(x, GpuSparseMatrix):
        raise TypeError("'x' must be a gpu_sparse_matrix object")
    lib = _bind(_lib.load())
    a, b, c, st = C.c_double(x.col_ptr), C.c_double(x.row_idx), C.c_double(x.values), C.c_int(0)
    lib.rcppml_sp_free_gpu(C.byref(a), C.byref(b), C.byref(c), C.byref(st))
    x.col_ptr = x.row_idx = x.values = 0.0
    return None


def nmf_zerocopy(gpu_mat: GpuSparseMatrix, k: int, *, maxit=100, tol=1e-4, seed=42, w_init=None, **kw):
    """`.gpu_nmf_zerocopy(gpu_mat, k, ...)` (R/gpu_backend.R:183-265): NMF on a device-resident matrix. W (k x m) comes
    from `w_init` (m x k) or from uniform draws of `seed`, H (k x n) from uniform draws (R's `runif` stream is not
    reproducible outside R; numpy's generator stands in for it). Remaining keywords: bridge.gpu_nmf_zerocopy."""
    from .bridge import gpu_nmf_zerocopy
    if not isinstance(gpu_mat, GpuSparseMatrix):
        raise TypeError("gpu_mat must be a gpu_sparse_matrix from sp_read_gpu()")
    rng = np.random.default_rng(seed)
    W_T0 = rng.random((gpu_mat.m, k)) if w_init is None else np.asarray(w_init, dtype=np.float64)
    H0 = rng.random((gpu_mat.n, k))
    return gpu_nmf_zerocopy(gpu_mat.col_ptr, gpu_mat.row_idx, gpu_mat.values, gpu_mat.m, gpu_mat.n, gpu_mat.nnz, k,
                            W_T0, H0, maxit=maxit, tol=tol, seed=seed, **kw)
