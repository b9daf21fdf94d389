"""Device-resident engine wrapper (ABI extension rcppml_b200_*, include/rcppml_gpu.h part 2).

Dense factors cross this boundary as C-contiguous float32 arrays of shape (cols, k): row c is
column c of the reference's k × cols column-major matrix (W_T: k × m, H: k × n).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Config, Result


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def make_config(k, *, max_iter=100, tol=1e-4, L1=(0.0, 0.0), L2=(0.0, 0.0), upper_bound=(0.0, 0.0),
                nonneg=(True, True), cd_maxit=100, cd_tol=1e-8, norm_type=0, solver_mode=0, patience=5,
                verbose=False) -> Config:
    """Pairs are (W, H) as at the R boundary (src/RcppFunctions_nmf.cpp:59-72)."""
    return Config(k=k, max_iter=max_iter, tol=tol, L1_W=L1[0], L1_H=L1[1], L2_W=L2[0], L2_H=L2[1],
                  ub_W=upper_bound[0], ub_H=upper_bound[1], nonneg_W=int(nonneg[0]), nonneg_H=int(nonneg[1]),
                  cd_maxit=cd_maxit, cd_tol=cd_tol, norm_type=norm_type, solver_mode=solver_mode,
                  patience=patience, verbose=int(verbose))


@dataclass
class FitResult:
    iterations: int
    converged: bool
    train_loss: float
    final_tol: float
    status: int
    gpu_launches: int
    loop_ms: float


class Engine:
    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.rcppml_b200_engine_create(C.byref(self._h), device), "engine_create")
        self.m = self.n = self.k = 0
        self.nnz = 0                 # non-zeros of this rank's column block
        self.nnz_global = 0
        self.col_begin = self.n_loc = self.row_begin = self.m_loc = 0

    def close(self):
        if self._h:
            self._lib.rcppml_b200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- matrix
    def set_matrix(self, m, n, indptr, indices, data):
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        nnz = int(indptr[n])
        if data.dtype == np.float64:
            data = np.ascontiguousarray(data)
            rc = self._lib.rcppml_b200_set_matrix_f64(self._h, m, n, nnz, _p(indptr, C.c_int), _p(indices, C.c_int),
                                                      _p(data, C.c_double))
        else:
            data = np.ascontiguousarray(data, dtype=np.float32)
            rc = self._lib.rcppml_b200_set_matrix_f32(self._h, m, n, nnz, _p(indptr, C.c_int), _p(indices, C.c_int),
                                                      _p(data, C.c_float))
        _lib.check(rc, "set_matrix")
        self._refresh_shard(m, n)

    def set_matrix_with_transpose(self, m, n, A, At):
        """A = (indptr, indices, data) of the m x n matrix, At = the same of its transpose (n x m, ascending column ids
        per row — e.g. a .spz file written with include_transpose): both are uploaded, no device transpose."""
        ap, ai = np.ascontiguousarray(A[0], np.int32), np.ascontiguousarray(A[1], np.int32)
        tp, ti = np.ascontiguousarray(At[0], np.int32), np.ascontiguousarray(At[1], np.int32)
        nnz = int(ap[n])
        if A[2].dtype == np.float64:
            ax, tx = np.ascontiguousarray(A[2], np.float64), np.ascontiguousarray(At[2], np.float64)
            rc = self._lib.rcppml_b200_set_matrix_with_transpose_f64(self._h, m, n, nnz, _p(ap, C.c_int), _p(ai, C.c_int),
                                                                     _p(ax, C.c_double), _p(tp, C.c_int), _p(ti, C.c_int),
                                                                     _p(tx, C.c_double))
        else:
            ax, tx = np.ascontiguousarray(A[2], np.float32), np.ascontiguousarray(At[2], np.float32)
            rc = self._lib.rcppml_b200_set_matrix_with_transpose_f32(self._h, m, n, nnz, _p(ap, C.c_int), _p(ai, C.c_int),
                                                                     _p(ax, C.c_float), _p(tp, C.c_int), _p(ti, C.c_int),
                                                                     _p(tx, C.c_float))
        _lib.check(rc, "set_matrix_with_transpose")
        self._refresh_shard(m, n)

    def set_matrix_spz(self, path, threads=0, stored_transpose=None):
        """A StreamPress v2 `.spz` file straight into the engine. After comm_init (+ set_partition) this
        rank decodes only its column block of A and its row block (columns of the file's transpose section) —
        collective, like every sharded set_matrix_*. stored_transpose: True use the file's transpose section when it is
        usable, False never, None automatic (sharded: yes; one GPU: no — the device transpose is ~100x cheaper than
        entropy-decoding the section). Returns True when the stored transpose was used."""
        from .streampress import SpzFile
        with SpzFile(path) as f:
            used = C.c_int(0)
            mode = -1 if stored_transpose is None else int(bool(stored_transpose))
            rc = self._lib.rcppml_b200_set_matrix_spz(self._h, f._h, int(threads), mode, C.byref(used))
            _lib.check(rc, "set_matrix_spz")
            self._refresh_shard(f.raw.m, f.raw.n)
        return bool(used.value)

    def set_matrix_synthetic(self, m, n_local, col_begin, density, seed):
        _lib.check(self._lib.rcppml_b200_set_matrix_synthetic(self._h, m, n_local, col_begin, density, seed),
                   "set_matrix_synthetic")
        self._refresh_shard(m, n_local)

    def _refresh_shard(self, m, n):
        cb, nl, rb, ml, ng = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int64(0)
        _lib.check(self._lib.rcppml_b200_get_shard(self._h, C.byref(cb), C.byref(nl), C.byref(rb), C.byref(ml),
                                                   C.byref(ng)), "get_shard")
        self.m, self.n = m, n
        self.col_begin, self.n_loc, self.row_begin, self.m_loc, self.nnz_global = (cb.value, nl.value, rb.value,
                                                                                  ml.value, ng.value)
        nnz = C.c_int64(0)
        _lib.check(self._lib.rcppml_b200_get_matrix(self._h, C.byref(nnz), None, None, None), "get_matrix")
        self.nnz = nnz.value

    def set_matrix_synthetic_sharded(self, m, n, density, seed):
        """This rank's column and row block of the m x n generator matrix (any world size)."""
        _lib.check(self._lib.rcppml_b200_set_matrix_synthetic_sharded(self._h, m, n, density, seed),
                   "set_matrix_synthetic_sharded")
        self._refresh_shard(m, n)

    def set_matrix_sharded(self, m, n, col_block, row_block):
        """col_block / row_block: (indptr, indices, data) as produced by rcppml_b200.shard.extract_*."""
        arrs = []
        for (p_, i_, x_) in (col_block, row_block):
            arrs += [np.ascontiguousarray(p_, np.int32), np.ascontiguousarray(i_, np.int32),
                     np.ascontiguousarray(x_, np.float32)]
        _lib.check(self._lib.rcppml_b200_set_matrix_sharded_f32(
            self._h, m, n, _p(arrs[0], C.c_int), _p(arrs[1], C.c_int), _p(arrs[2], C.c_float),
            _p(arrs[3], C.c_int), _p(arrs[4], C.c_int), _p(arrs[5], C.c_float)), "set_matrix_sharded")
        self._refresh_shard(m, n)

    def set_partition(self, col_cuts=None, row_cuts=None):
        """Explicit partition for the following set_matrix_* calls (same cuts on every rank): world+1 ascending cuts
        0 .. n / 0 .. m (rcppml_b200.shard.balanced_cuts); None restores equal blocks."""
        cc = None if col_cuts is None else np.ascontiguousarray(col_cuts, np.int32)
        rc = None if row_cuts is None else np.ascontiguousarray(row_cuts, np.int32)
        _lib.check(self._lib.rcppml_b200_set_partition(self._h, _p(cc, C.c_int), _p(rc, C.c_int)), "set_partition")

    def factor_checksum(self):
        """64-bit device-side checksums (W_T, H, d) of the logical factors: equal <=> bit-identical."""
        out = (C.c_uint64 * 3)()
        _lib.check(self._lib.rcppml_b200_factor_checksum(self._h, out), "factor_checksum")
        return int(out[0]), int(out[1]), int(out[2])

    def get_matrix(self):
        p = np.empty(self.n_loc + 1, np.int32)
        i = np.empty(self.nnz, np.int32)
        x = np.empty(self.nnz, np.float32)
        nnz = C.c_int64(0)
        _lib.check(self._lib.rcppml_b200_get_matrix(self._h, C.byref(nnz), _p(p, C.c_int), _p(i, C.c_int),
                                                    _p(x, C.c_float)), "get_matrix")
        return p, i, x

    def get_matrix_t(self):
        p = np.empty(self.m_loc + 1, np.int32)
        _lib.check(self._lib.rcppml_b200_get_matrix_t(self._h, _p(p, C.c_int), None, None), "get_matrix_t")
        cnt = int(p[-1])
        i = np.empty(cnt, np.int32)
        x = np.empty(cnt, np.float32)
        _lib.check(self._lib.rcppml_b200_get_matrix_t(self._h, None, _p(i, C.c_int), _p(x, C.c_float)),
                   "get_matrix_t")
        return p, i, x

    def set_mask(self, mask_indptr=None, mask_indices=None):
        """CSC pattern (m x n) of masked entries (nmf/masked_nnls.hpp); None clears the mask."""
        if mask_indptr is None:
            _lib.check(self._lib.rcppml_b200_set_mask(self._h, 0, None, None), "set_mask")
            return
        mp = np.ascontiguousarray(mask_indptr, np.int32)
        mi = np.ascontiguousarray(mask_indices, np.int32)
        _lib.check(self._lib.rcppml_b200_set_mask(self._h, int(mp[self.n]), _p(mp, C.c_int), _p(mi, C.c_int)),
                   "set_mask")

    # ---- factors
    def set_factors(self, W_T, H):
        W_T = np.ascontiguousarray(W_T, dtype=np.float32)
        H = np.ascontiguousarray(H, dtype=np.float32)
        if W_T.ndim != 2 or H.ndim != 2 or W_T.shape[0] != self.m or H.shape[0] != self.n or W_T.shape[1] != H.shape[1]:
            raise ValueError(f"set_factors: expected W_T ({self.m}, k) and H ({self.n}, k)")
        self.k = W_T.shape[1]
        _lib.check(self._lib.rcppml_b200_set_factors_f32(self._h, self.k, _p(W_T, C.c_float), _p(H, C.c_float)),
                   "set_factors")

    def init_factors(self, k, seed, h_col_begin=0):
        _lib.check(self._lib.rcppml_b200_init_factors(self._h, k, seed, h_col_begin), "init_factors")
        self.k = k

    def get_factors(self, out=None):
        """Returns (W_T (m, k), H (n, k), d (k)) as float32. `out`: optional preallocated (W_T, H, d) C-contiguous
        float32 arrays of those shapes (e.g. pinned memory) that receive the copies instead of fresh arrays."""
        if out is not None:
            W_T, H, d = out
            for a, shape in ((W_T, (self.m, self.k)), (H, (self.n, self.k)), (d, (self.k,))):
                if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == shape
                        and a.flags.c_contiguous and a.flags.writeable):
                    raise ValueError(f"get_factors(out=...): expected a writeable C-contiguous float32 array of shape {shape}")
        else:
            W_T = np.empty((self.m, self.k), np.float32)
            H = np.empty((self.n, self.k), np.float32)
            d = np.empty(self.k, np.float32)
        _lib.check(self._lib.rcppml_b200_get_factors_f32(self._h, _p(W_T, C.c_float), _p(H, C.c_float),
                                                         _p(d, C.c_float)), "get_factors")
        return W_T, H, d

    def set_factor_blocks(self, W_blk, H_blk):
        """Sharded fits: upload only this rank's rows of W_T ((m_loc, k), rows row_begin..) and of H ((n_loc, k),
        columns col_begin..); the library completes the replicas with one all-gather per factor over NVLink.
        Collective: every rank must call it."""
        W_blk = np.ascontiguousarray(W_blk, dtype=np.float32)
        H_blk = np.ascontiguousarray(H_blk, dtype=np.float32)
        if W_blk.shape[0] != self.m_loc or H_blk.shape[0] != self.n_loc or W_blk.shape[1] != H_blk.shape[1]:
            raise ValueError(f"set_factor_blocks: expected ({self.m_loc}, k) and ({self.n_loc}, k)")
        self.k = W_blk.shape[1]
        _lib.check(self._lib.rcppml_b200_set_factor_blocks_f32(self._h, self.k, _p(W_blk, C.c_float),
                                                               _p(H_blk, C.c_float)), "set_factor_blocks")

    def get_factor_blocks(self, out=None):
        """This rank's blocks of the fitted factors: (W_T[row_begin : row_begin + m_loc], H[col_begin : col_begin +
        n_loc], d). `out` as in get_factors."""
        shapes = ((self.m_loc, self.k), (self.n_loc, self.k), (self.k,))
        if out is not None:
            for a, shape in zip(out, shapes):
                if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == shape
                        and a.flags.c_contiguous and a.flags.writeable):
                    raise ValueError(f"get_factor_blocks(out=...): expected a writeable C-contiguous float32 array of shape {shape}")
            W_blk, H_blk, d = out
        else:
            W_blk, H_blk, d = (np.empty(s_, np.float32) for s_ in shapes)
        _lib.check(self._lib.rcppml_b200_get_factor_blocks_f32(self._h, _p(W_blk, C.c_float), _p(H_blk, C.c_float),
                                                               _p(d, C.c_float)), "get_factor_blocks")
        return W_blk, H_blk, d

    # ---- fit
    def begin_fit(self, cfg: Config):
        _lib.check(self._lib.rcppml_b200_begin_fit(self._h, C.byref(cfg)), "begin_fit")

    def iterate(self, n_iters: int):
        _lib.check(self._lib.rcppml_b200_iterate(self._h, n_iters), "iterate")

    def fit(self, cfg: Config) -> FitResult:
        _lib.check(self._lib.rcppml_b200_fit(self._h, C.byref(cfg)), "fit")
        return self.result()

    def fit_cv(self, cfg: Config, *, holdout_fraction=0.1, cv_seed=0, seed=42, mask_zeros=True, cv_patience=5):
        """nmf_fit_cv (nmf/fit_cv.hpp:124): speckled-mask cross-validation. Returns (FitResult, cv dict)."""
        cv = _lib.CvConfig(holdout_fraction=holdout_fraction, cv_seed=cv_seed, seed=seed, mask_zeros=int(mask_zeros),
                           cv_patience=cv_patience)
        _lib.check(self._lib.rcppml_b200_fit_cv(self._h, C.byref(cfg), C.byref(cv)), "fit_cv")
        r = _lib.CvResult()
        _lib.check(self._lib.rcppml_b200_get_cv_result(self._h, C.byref(r)), "get_cv_result")
        res = self.result()
        it = res.iterations
        tr, te = np.zeros(max(it, 1), np.float32), np.zeros(max(it, 1), np.float32)
        _lib.check(self._lib.rcppml_b200_get_cv_history(self._h, _p(tr, C.c_float), _p(te, C.c_float), it), "cv_history")
        return res, dict(train_loss=r.train_loss, test_loss=r.test_loss, best_test_loss=r.best_test_loss,
                         best_iter=r.best_iter, n_test=r.n_test, train_history=tr[:it], test_history=te[:it])

    def half_step(self, cfg: Config, which: int, warm_start: bool, normalize_after: bool = False):
        _lib.check(self._lib.rcppml_b200_half_step(self._h, C.byref(cfg), which, int(warm_start),
                                                   int(normalize_after)), "half_step")

    def result(self) -> FitResult:
        r = Result()
        _lib.check(self._lib.rcppml_b200_get_result(self._h, C.byref(r)), "get_result")
        return FitResult(r.iterations, bool(r.converged), r.train_loss, r.final_tol, r.status, r.gpu_launches,
                         r.loop_ms)

    def loss_history(self, count: int) -> np.ndarray:
        out = np.zeros(max(count, 1), np.float32)
        _lib.check(self._lib.rcppml_b200_get_loss_history(self._h, _p(out, C.c_float), count), "loss_history")
        return out[:count]

    def cd_sweeps(self) -> int:
        return int(self._lib.rcppml_b200_cd_sweeps(self._h))

    def set_profiling(self, on: bool):
        _lib.check(self._lib.rcppml_b200_set_profiling(self._h, int(on)), "set_profiling")

    def profile(self):
        ms = (C.c_double * _lib.NUM_SECTIONS)()
        ln = (C.c_int * _lib.NUM_SECTIONS)()
        _lib.check(self._lib.rcppml_b200_get_profile(self._h, ms, ln), "get_profile")
        return ({name: ms[i] for i, name in enumerate(_lib.SECTION_NAMES)},
                {name: ln[i] for i, name in enumerate(_lib.SECTION_NAMES)})

    def counters(self):
        a, b = C.c_int64(0), C.c_int64(0)
        _lib.check(self._lib.rcppml_b200_get_counters(self._h, C.byref(a), C.byref(b)), "get_counters")
        return a.value, b.value

    # ---- multi-GPU
    def comm_init(self, rank: int, world: int, unique_id: bytes):
        _lib.check(self._lib.rcppml_b200_comm_init(self._h, rank, world, unique_id), "comm_init")

    def comm_enable_p2p(self, dist) -> bool:
        """Peer-memory fast path (include/rcppml_gpu.h). `dist` is an initialised torch.distributed (only used to
        all-gather the per-rank handle blobs). Call on every rank after the factors exist. Two flavours, chosen by the
        library at comm_init: NVSwitch MULTICAST (comm_mc_export/import: the factors are VMM allocations bound to
        multicast objects; a normalised block is written into every replica with one multimem.st per word) where the
        devices support it and RCPPML_B200_MC != 0, else unicast peer stores over CUDA IPC mappings
        (comm_ipc_export/import) — also the fall-back when the multicast set-up fails on any rank. Returns False (and
        leaves the NCCL loop active) when RCPPML_B200_P2P=0."""
        import os
        import torch
        self.p2p_mode = "nccl"
        if os.environ.get("RCPPML_B200_P2P", "1") == "0":
            return False
        world = dist.get_world_size()
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"

        def all_ok(rc):
            t = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(t.item() == 1)

        def gather(raw):
            mine = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            allb = torch.empty(world * len(raw), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allb, mine)
            return allb.cpu().numpy().tobytes()

        if self._lib.rcppml_b200_comm_mc_wanted(self._h):
            buf = C.create_string_buffer(128)
            ok = all_ok(self._lib.rcppml_b200_comm_mc_export(self._h, buf))
            if ok:
                ok = all_ok(self._lib.rcppml_b200_comm_mc_import(self._h, gather(buf.raw)))
            if ok:                                           # every device joined: binding cannot block any more
                ok = all_ok(self._lib.rcppml_b200_comm_mc_bind(self._h))
            dist.barrier()                                   # every rank duplicated the descriptors it needs
            self._lib.rcppml_b200_comm_mc_finish(self._h)
            if ok:
                self.p2p_mode = "multicast"
                return True
            # fall back to unicast peer stores: the factors move into plain allocations that CUDA IPC can export
            if not all_ok(self._lib.rcppml_b200_comm_mc_disable(self._h)):
                self._lib.rcppml_b200_comm_p2p_close(self._h)
                return False
        buf = C.create_string_buffer(192)
        _lib.check(self._lib.rcppml_b200_comm_ipc_export(self._h, buf), "comm_ipc_export")
        _lib.check(self._lib.rcppml_b200_comm_ipc_import(self._h, gather(buf.raw)), "comm_ipc_import")
        dist.barrier()
        self.p2p_mode = "unicast"
        return True


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _lib.check(_lib.load().rcppml_b200_nccl_unique_id(buf), "nccl_unique_id")
    return buf.raw
