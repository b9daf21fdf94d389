"""Host mirror of the reference's `nmf()` front end for the path this repository replaces.

Argument names, defaults, meaning and error messages follow R/nmf_thin.R:219-1315 (`nmf`), R/parse_dots.R:5-96
(`.parse_nmf_dots`) and R/nmf_validation.R:64-240; the R -> config mapping follows
src/RcppFunctions_nmf.cpp:23-95, 335-455 (`build_config_from_params`, `Rcpp_nmf_full`: data and W_init are cast to
float) and the call reaches the GPU exactly as `nmf::nmf` does with plan = GPU (nmf/fit.hpp:97-117, 187-203):
through the dlsym bridge's packing (gpu/bridge_nmf.hpp:199-393; rcppml_b200/bridge.py is its ctypes twin) into
`rcppml_gpu_nmf_unified_float` / `rcppml_gpu_nmf_cv_unified_float` of RcppML_gpu.so.

What is mirrored: sparse `data`, scalar `k`, `tol`, `maxit`, `L1`, `L2`, `seed` (NULL / integer / W matrix / vector of
seeds or list of W matrices = several initialisations, best loss kept),
`mask` (NULL / "zeros" / pattern matrix / list("zeros", matrix)), `nonneg`, `test_fraction`, and from `...`:
`upper_bound`, `norm`, `solver`, `cd_maxit`, `cd_tol`, `h_init`, `w_init`, `cv_seed`, `patience`, `sort_model`,
`resource`, `threads`. Everything the reference routes elsewhere (non-MSE losses, robust/zi, L21 / angular / graph /
target penalties, projective / symmetric, rank vectors and k = "auto", SVD-based init strings, streaming .spz,
dense input) raises NotImplementedError naming the argument: SURVEY.md §2 marks them out of
scope and there is NO CPU fallback behind this function — where the reference would fall back to its CPU loop
(nmf/fit.hpp:118-127), this one raises `NativeLibraryError`.

Two points where the GPU route of the reference loses information and this mirror does not:
  * an explicit `mask` is not carried by the bridge signature — here it goes to the ABI
    extension `rcppml_gpu_nmf_masked_unified_float`;
  * `sort_model` is not transmitted either (core/config.hpp:358), so the reference returns unsorted factors from
    the GPU route; here `sort_model=True` (the R default) applies `NMFResult::sort` (core/result.hpp:169-189) on
    the host, which is what the reference's CPU route returns.
"""
from __future__ import annotations

import time
import warnings
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .bridge import bridge_nmf_cv_sparse, bridge_nmf_sparse, gpu_detect

_INT_MAX = 2147483647          # .Machine$integer.max

# .parse_nmf_dots defaults (R/parse_dots.R:6-58)
_DOT_DEFAULTS = dict(
    L21=(0, 0), angular=(0, 0), upper_bound=(0, 0), graph_W=None, graph_H=None, graph_lambda=(0, 0),
    target_H=None, target_lambda=(0, 0), dispersion="per_row", theta_init=0.1, theta_max=5.0, theta_min=0.0,
    nb_size_init=10.0, nb_size_max=1e6, nb_size_min=0.01, gamma_phi_init=1.0, gamma_phi_max=1e4,
    gamma_phi_min=1e-6, tweedie_power=1.5, irls_max_iter=5, irls_tol=1e-4, huber_delta=1.0, zi_em_iters=1,
    solver="auto", cd_tol=1e-8, cd_maxit=100, h_init=None, w_init=None, cv_seed=None, patience=5,
    cv_k_range=(2, 50), track_train_loss=True, threads=0, resource="auto", norm="L1", sort_model=True,
    streaming="auto", panel_cols=0, dispatch=None, on_iteration=None, profile=False, convergence="loss",
    sparse=False)


# ------------------------------------------------------------------------------------------------ R's RNG ----
class RRandom:
    """R's default generator (Mersenne-Twister + inversion), enough of it for `set.seed(s); runif(n)` — what
    nmf() draws W from (R/nmf_thin.R:794-795). R-4 src/main/RNG.c: `RNG_Init` scrambles the seed with
    `seed = 69069 * seed + 1` 50 times, fills dummy[0..624] with the next 625 values of that LCG, `FixupSeeds`
    sets dummy[0] = mti = 624; `MT_genrand` is the standard MT19937 tempering, scaled by 2.3283064365386963e-10
    and clamped into (0, 1) by `fixup`. Pinned by the well-known values of set.seed(42) / (1) / (123) in
    tests/test_nmf_api.py."""

    _I2_32M1 = 2.328306437080797e-10

    def __init__(self, seed: int):
        s = int(seed) & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        words = np.empty(625, np.uint32)
        for j in range(625):
            s = (69069 * s + 1) & 0xFFFFFFFF
            words[j] = s
        self._bg = np.random.MT19937()
        st = self._bg.state
        st["state"]["key"] = words[1:].copy()
        st["state"]["pos"] = 624
        self._bg.state = st

    def runif(self, n: int) -> np.ndarray:
        u = self._bg.random_raw(int(n)).astype(np.float64) * 2.3283064365386963e-10
        u[u <= 0.0] = 0.5 * self._I2_32M1
        u[(1.0 - u) <= 0.0] = 1.0 - 0.5 * self._I2_32M1
        return u


# ------------------------------------------------------------------------------------ SplitMix64 (rng.hpp) ----
_GAMMA = 0x9E3779B97F4A7C15


def splitmix_fill_uniform_f64(seed: int, count: int) -> np.ndarray:
    """`SplitMix64(seed).fill_uniform(double*, ...)` (rng/rng.hpp:73, 89-104, 195-201): element e of the stream is
    mix(state0 + (e+1)·γ); uniform<double>() = double(next()) / double(UINT64_MAX) (= 2^64 after rounding)."""
    s0 = 12345 if (seed & 0xFFFFFFFFFFFFFFFF) == 0 else (seed & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        z = np.uint64(s0) + np.arange(1, count + 1, dtype=np.uint64) * np.uint64(_GAMMA)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.astype(np.float64) / 18446744073709551616.0


def bridge_h_init(seed: int, k: int, n: int) -> np.ndarray:
    """H as the bridge draws it when no h_init is given (gpu/bridge_nmf.hpp:226-229): a fresh double stream seeded
    `config.seed + 0x9E3779B9u` (32-bit unsigned wrap). Returned as (n, k): row j = column j of the k x n H."""
    return splitmix_fill_uniform_f64((int(seed) + 0x9E3779B9) & 0xFFFFFFFF, k * n).reshape(n, k)


def bridge_w_init(seed: int, k: int, m: int) -> np.ndarray:
    """W_T when neither seed nor w_init yields a matrix (gpu/bridge_nmf.hpp:216-218). R always sends one, so this is
    reached only with an SVD-init string, which this mirror refuses; kept for direct callers."""
    return splitmix_fill_uniform_f64(int(seed) & 0xFFFFFFFF, k * m).reshape(m, k)


# -------------------------------------------------------------------------------------------- validation ----
class _SeedOnly(int):
    """config.seed for one run of a multi-initialisation fit: the W matrix comes through `w_init`, so the scalar-seed
    branch must not draw one (R/nmf_thin.R:839-853: `seed = seed_int + i - 1L, w_init_sexp = w_init_list[[i]]`)."""


def _pair(value, name):
    """validate_penalty (R/nmf_validation.R:64-76)."""
    v = np.atleast_1d(np.asarray(value, dtype=np.float64))
    if v.size == 1:
        v = np.repeat(v, 2)
    elif v.size != 2:
        raise ValueError(f"'{name}' must be length 1 or 2 for c(w, h)")
    if (v < 0).any():
        raise ValueError(f"'{name}' values must be non-negative")
    return float(v[0]), float(v[1])


def _parse_dots(dots: dict) -> dict:
    unknown = [k for k in dots if k not in _DOT_DEFAULTS]
    if unknown:
        raise TypeError("Unknown parameter(s) passed to nmf(): " + ", ".join(f"'{u}'" for u in unknown)
                        + ". See ?nmf for valid parameters.")
    out = dict(_DOT_DEFAULTS)
    out.update(dots)
    for name, choices in (("solver", ("auto", "cholesky", "cd")), ("norm", ("L1", "L2", "none"))):
        if out[name] not in choices:
            raise ValueError(f"'{name}' should be one of " + ", ".join(f"'{c}'" for c in choices))
    return out


def _refuse(name, why="outside the ALS hot path this library replaces (SURVEY.md §2); use the reference's CPU route"):
    raise NotImplementedError(f"nmf(): '{name}' is {why}")


def _as_csc(data):
    import scipy.sparse as sp
    if isinstance(data, str):
        _refuse("data = <file path>", "file input (.spz streaming and loaders) is outside this library's scope")
    if isinstance(data, (list, tuple, dict)):
        _refuse("data = <list>", "multi-modal input (factor_net) is outside this library's scope")
    if not sp.issparse(data):
        arr = np.asarray(data)
        if arr.ndim != 2 or not np.issubdtype(arr.dtype, np.number):
            raise ValueError("'data' was not coercible to a numeric matrix")
        _refuse("dense data", "the dense-A path (rcppml_gpu_nmf_dense_unified_float) is out of scope; pass a "
                              "scipy.sparse matrix (the dgCMatrix route)")
    A = data.tocsc().astype(np.float32)          # Rcpp_nmf_full: A_mapped.cast<float>()
    A.sort_indices()
    if np.isnan(A.data).any():
        _refuse("NA values in data", "the NA-mask route is not mirrored; pass an explicit mask")
    return A


def _mask_pattern(mask, shape):
    """validate_mask (R/nmf_validation.R:173-216) -> (pattern CSC or None, mask_zeros)."""
    import scipy.sparse as sp
    if mask is None:
        return None, False
    if isinstance(mask, (list, tuple)) and not sp.issparse(mask):
        if len(mask) < 2 or not isinstance(mask[0], str) or mask[0] != "zeros":
            raise ValueError("'mask' list must be list(\"zeros\", <matrix>)")
        pat, _ = _mask_pattern(mask[1], shape)
        return pat, True
    if isinstance(mask, str):
        if mask == "zeros":
            return None, True
        if mask == "NA":
            return None, False
        raise ValueError("'mask' must be NULL, 'zeros', 'NA', a matrix, or list(\"zeros\", <matrix>)")
    try:
        M = sp.csc_matrix(mask)
    except Exception as exc:        # noqa: BLE001
        raise ValueError("could not coerce the value of 'mask' to a sparse pattern matrix (dgCMatrix)") from exc
    if M.shape != tuple(shape):
        raise ValueError("'mask' dimensions must match 'data'")
    M.eliminate_zeros()
    M.sort_indices()
    return (M if M.nnz > 0 else None), False          # has_mask() is false for an empty matrix (core/config.hpp:254)


def sort_by_d(w, d, h):
    """NMFResult::sort (core/result.hpp:169-189): factors in decreasing order of d."""
    order = np.argsort(-np.asarray(d, dtype=np.float64), kind="stable")
    return w[:, order], d[order], h[order, :]


# ------------------------------------------------------------------------------------------------- result ----
@dataclass
class NMFModel:
    """The reference's S4 class `nmf`: w (m x k), d (k), h (k x n), misc (R/nmf_thin.R:1225-1313)."""
    w: np.ndarray
    d: np.ndarray
    h: np.ndarray
    misc: dict = field(default_factory=dict)

    def __iter__(self):                      # w, d, h = model
        return iter((self.w, self.d, self.h))

    def reconstruct(self):
        return (self.w * self.d) @ self.h

    def predict(self, data, L1=None, L2=None, mask=None, upper_bound=None, threads=0, verbose=False):
        """`predict(object, data, ...)` (R/predict_nmf.R:48-98): project new columns onto w (GPU, fp64). Penalties
        default to the model's own h-side values; returns a new model with the projected h, misc = {projected}.
        `mask` is validated and then unused, as in the reference (Rcpp_predict never reads it)."""
        from .project import predict
        import scipy.sparse as sp
        if L1 is None:
            L1 = self.misc["L1"][1] if self.misc.get("L1") is not None else 0
        if L2 is None:
            L2 = self.misc["L2"][1] if self.misc.get("L2") is not None else 0
        if upper_bound is None:
            upper_bound = self.misc["upper_bound"][1] if self.misc.get("upper_bound") is not None else 0
        if np.size(L1) != 1:
            raise ValueError("'L1' must be a single value giving the penalty on 'h'")
        if L1 >= 1 or L1 < 0:
            raise ValueError("L1 penalty must be strictly in the range [0,1)")
        if np.size(L2) != 1:
            raise ValueError("'L2' must be a single value giving the penalty on 'h'")
        if L2 < 0:
            raise ValueError("L2 penalty must be strictly >= 0")
        A = data.tocsc() if sp.issparse(data) else sp.csc_matrix(np.asarray(data, np.float64))
        _mask_pattern(mask, A.shape)
        w = self.w
        if not (w.shape[0] == A.shape[0]) and w.shape[1] == A.shape[0]:
            w = w.T
        if w.shape[0] != A.shape[0]:
            raise ValueError("dimensions of 'object@w' and 'A' are not compatible")
        h = predict(w, A, L1=float(L1), L2=float(L2), upper_bound=float(upper_bound))
        return NMFModel(self.w, self.d, h, dict(projected=True))

    def evaluate(self, data, mask=None, missing_only=False, loss="mse", test_fraction=0, test_seed=None, eval_set="all",
                 threads=0, verbose=False):
        """`evaluate(x, data, mask, ...)` (R/nmf_methods.R:332-440 -> Rcpp_evaluate_loss) with loss = "mse": mean
        squared error over all entries, or over the non-zeros when mask = "zeros" (GPU, fp64, no dense m x n
        reconstruction). Explicit masks, missing_only and test-set evaluation are not mirrored."""
        from .project import evaluate
        if eval_set not in ("all", "test", "train"):
            raise ValueError("'eval_set' should be one of 'all', 'test', 'train'")
        if missing_only and mask is None:
            raise ValueError("a mask matrix must be specified to set 'missing_only = TRUE'")
        if loss not in ("mse", "gp"):
            raise ValueError("'loss' should be one of 'mse', 'gp'")
        if loss != "mse":
            _refuse("evaluate(loss = 'gp')")
        if test_fraction > 0 or test_seed is not None or eval_set != "all":
            _refuse("evaluate(test_fraction / test_seed / eval_set)", "not mirrored")
        if mask is not None and not (isinstance(mask, str) and mask == "zeros"):
            _refuse("evaluate(mask = <matrix>)", "not mirrored; only NULL and \"zeros\"")
        return evaluate(data, self.w, self.d, self.h, mask_zeros=(mask is not None))


def _multi_init(w_init_list, seed_int, start, test_fraction, call_args):
    """Multiple initialisations: run each, keep the lowest loss (R/nmf_thin.R:829-925). Run i (0-based) gets
    config.seed = seed_int + i — hence its own H stream — and w_init_list[i]."""
    if test_fraction > 0:
        raise ValueError("Multiple initializations are not compatible with cross-validation. Use a single seed or matrix.")
    args = dict(call_args)
    args.pop("w_init", None)
    best, best_idx, losses = None, 0, []
    for i, w0 in enumerate(w_init_list):
        model = nmf(seed=_SeedOnly(seed_int + i), w_init=w0, **args)
        losses.append(model.misc["loss"])
        if best is None or model.misc["loss"] < best.misc["loss"]:
            best, best_idx = model, i
    best.misc = dict(tol=best.misc["tol"], iter=best.misc["iter"], loss_type=best.misc["loss_type"], loss=best.misc["loss"],
                     runtime=time.time() - start, w_init=w_init_list[best_idx],
                     all_inits=[dict(init=i + 1, loss=v, selected=(i == best_idx)) for i, v in enumerate(losses)])
    return best


# --------------------------------------------------------------------------------------------------- nmf ----
def nmf(data, k, tol=1e-4, maxit=100, L1=(0, 0), L2=(0, 0), seed=None, mask=None, loss="mse", nonneg=(True, True),
        test_fraction=0, verbose=False, projective=False, symmetric=False, zi="none", robust=False, **dots) -> NMFModel:
    """Non-negative matrix factorisation A ~ w diag(d) h on the GPU (see the module docstring for the mapping)."""
    start = time.time()
    p = _parse_dots(dots)
    call_args = dict(data=data, k=k, tol=tol, maxit=maxit, L1=L1, L2=L2, mask=mask, loss=loss, nonneg=nonneg,
                     verbose=False, projective=projective, symmetric=symmetric, zi=zi, robust=robust, **dots)

    # ---- arguments whose code paths live outside the hot path (R/nmf_thin.R:278-420) ----
    if loss != "mse":
        if loss not in ("gp", "nb", "gamma", "inverse_gaussian", "tweedie"):
            raise ValueError("'loss' should be one of 'mse', 'gp', 'nb', 'gamma', 'inverse_gaussian', 'tweedie'")
        _refuse(f"loss = '{loss}'")
    if zi != "none":
        if zi not in ("row", "col"):
            raise ValueError("'zi' should be one of 'none', 'row', 'col'")
        raise ValueError("zi != 'none' requires loss='gp' or loss='nb'.")
    if robust is not False and robust != 0:
        _refuse("robust")
    if bool(projective) and bool(symmetric):
        raise ValueError("'projective' and 'symmetric' cannot both be TRUE")
    if projective:
        _refuse("projective")
    if symmetric:
        _refuse("symmetric")
    for name in ("graph_W", "graph_H", "target_H", "on_iteration", "dispatch"):
        if p[name] is not None:
            _refuse(name)
    if p["streaming"] is True or p["panel_cols"]:
        _refuse("streaming")
    if p["sparse"]:
        _refuse("sparse")
    if p["resource"] not in ("auto", "cpu", "gpu"):
        raise ValueError("'resource' must be \"auto\", \"cpu\", or \"gpu\"")
    if p["resource"] == "cpu":
        _refuse("resource = 'cpu'", "not available: this library is the GPU route and has no CPU fallback")

    A = _as_csc(data)
    m, n = A.shape

    # ---- penalties (validate_all_penalties, R/nmf_validation.R:78-111) ----
    L1 = _pair(L1, "L1")
    if max(L1) >= 1 or min(L1) < 0:
        raise ValueError("L1 penalties must be strictly in the range [0,1)")
    L2 = _pair(L2, "L2")
    if max(_pair(p["L21"], "L21")) > 0:
        _refuse("L21")
    if max(_pair(p["angular"], "angular")) > 0:
        _refuse("angular")
    _pair(p["graph_lambda"], "graph_lambda")
    upper_bound = _pair(p["upper_bound"], "upper_bound")

    # ---- validate_simple_params / validate_cv_params ----
    sort_model = p["sort_model"]
    if not isinstance(sort_model, (bool, np.bool_)):
        raise ValueError("'sort_model' must be a single logical value")
    nn = np.atleast_1d(np.asarray(nonneg))
    if nn.dtype != np.bool_:
        raise ValueError("'nonneg' must be logical")
    if nn.size == 1:
        nn = np.repeat(nn, 2)
    if nn.size != 2:
        raise ValueError("'nonneg' must be length 1 or 2 with no NA values")
    nonneg = (bool(nn[0]), bool(nn[1]))
    if not np.isscalar(test_fraction) or isinstance(test_fraction, (str, bool)):
        raise ValueError("'test_fraction' must be a single numeric value")
    if test_fraction < 0 or test_fraction >= 1:
        raise ValueError("'test_fraction' must be in the range [0, 1)")
    patience = p["patience"]
    if not np.isscalar(patience) or isinstance(patience, str):
        raise ValueError("'patience' must be a single numeric value")

    mask_pat, mask_zeros = _mask_pattern(mask, (m, n))

    # ---- k ----
    if isinstance(k, str):
        if k == "auto":
            _refuse("k = 'auto'", "rank search is outside this library's scope")
        raise ValueError("'k' must be a positive integer")
    if np.size(k) != 1:
        _refuse("k = <vector>", "multi-rank cross-validation sweeps are outside this library's scope")
    k = int(np.asarray(k).reshape(-1)[0])
    if k < 1:
        raise ValueError("'k' must be a positive integer")

    # ---- solver (R/nmf_thin.R:362-388; the GPU branch of "auto") ----
    solver = p["solver"]
    if solver == "auto":
        solver = "cd" if k <= 32 else "cholesky"
    solver_mode = {"cd": 0, "cholesky": 1}[solver]

    # ---- seed -> w_init + seed_int (R/nmf_thin.R:723-827) ----
    w_init = None
    seed_int = 0
    if isinstance(seed, str):
        if seed == "random":
            seed = None
        elif seed in ("lanczos", "irlba", "randomized", "svd"):
            _refuse(f"seed = '{seed}'", "SVD-based initialisation is outside this library's scope")
        else:
            raise ValueError(f"Unknown seed string '{seed}'. Valid: NULL, integer, matrix, "
                             "'lanczos', 'irlba', 'randomized', 'svd'")
    if seed is None:
        # R draws from the session's current RNG state; the host equivalent is numpy's global-free default_rng()
        w_init = np.random.default_rng().random((m, k))
        head = w_init.T.reshape(-1)[:min(10, w_init.size)]               # R's column-major w_init_mat[1:10]
        seed_int = int(abs(head.sum() * 1e8) % _INT_MAX)
    elif isinstance(seed, (list, tuple)) and len(seed) and all(isinstance(x, np.ndarray) for x in seed):
        w_init_list = []                                                  # list of custom W matrices (:745-757)
        for x in seed:
            if x.ndim != 2:
                raise ValueError("Each element of seed list must be a matrix")
            actual_k = x.shape[1] if x.shape[0] == m else x.shape[0]
            if actual_k != k:
                raise ValueError(f"Rank mismatch: k={k} specified but custom initialization has rank {actual_k}.")
            if x.shape[0] == m:
                w_init_list.append(np.asarray(x, np.float64))
            elif x.shape[1] == m:
                w_init_list.append(np.asarray(x, np.float64).T)
            else:
                raise ValueError("Custom init matrix dimensions incompatible with data")
        seed_int = abs(int(float(np.sum(seed[0] * 1e6)) % _INT_MAX))
        if len(w_init_list) > 1:
            return _multi_init(w_init_list, seed_int, start, test_fraction, call_args)
        w_init = bridge_w_init(seed_int, k, m)              # a one-element list leaves w_init_mat NULL: bridge_nmf.hpp:216
    elif isinstance(seed, np.ndarray) and seed.ndim == 2:
        actual_k = seed.shape[1] if seed.shape[0] == m else seed.shape[0]
        if actual_k != k:
            raise ValueError(f"Rank mismatch: k={k} specified but custom initialization has rank {actual_k}.")
        if seed.shape[0] == m:
            w_init = np.asarray(seed, np.float64)
        elif seed.shape[1] == m:
            w_init = np.asarray(seed, np.float64).T
        else:
            raise ValueError("Custom init matrix dimensions incompatible with data")
        seed_int = abs(int(float(np.sum(seed * 1e6)) % _INT_MAX))
    elif np.size(seed) > 1:
        svec = [int(x) for x in np.asarray(seed).reshape(-1)]            # vector of seeds for multi-init (:772-791)
        w_init_list = [RRandom(sv).runif(m * k).reshape(k, m).T for sv in svec]
        return _multi_init(w_init_list, svec[0], start, test_fraction, call_args)
    elif isinstance(seed, _SeedOnly):
        seed_int = int(seed)
    elif isinstance(seed, (int, float, np.integer, np.floating)):
        seed_int = int(seed)
        w_init = RRandom(seed_int).runif(m * k).reshape(k, m).T          # matrix(runif(m*k), m, k): column-major
    else:
        raise ValueError("'seed' must be NULL, an integer, or a matrix")

    if p["w_init"] is not None:
        wi = np.asarray(p["w_init"], np.float64)
        if wi.ndim != 2:
            raise ValueError("w_init must be a matrix")
        if wi.shape == (m, k):
            w_init = wi
        elif wi.shape == (k, m):
            w_init = wi.T
        else:
            raise ValueError("w_init dimensions incompatible with data and k")
        if seed_int == 0 and not isinstance(seed, _SeedOnly):
            head = w_init.T.reshape(-1)[:min(10, w_init.size)]
            seed_int = abs(int(head.sum() * 1e8) % _INT_MAX)

    if w_init is None:                                                    # no matrix from R: the bridge draws W (:216-218)
        w_init = bridge_w_init(seed_int, k, m)
    # Rcpp_nmf_full casts W_init / H_init to float before the bridge widens them again (:419-433)
    W_T0 = np.ascontiguousarray(w_init, dtype=np.float32)                 # (m, k): row = column of the k x m W_T
    if p["h_init"] is not None:
        hi = np.asarray(p["h_init"], np.float64)
        if hi.shape != (k, n):                                            # bridge_nmf.hpp:221 ignores other shapes
            raise ValueError("h_init must be a k x n matrix")
        H0 = np.ascontiguousarray(hi.T, dtype=np.float32)
    elif test_fraction > 0:
        H0 = None                                                         # the CV bridge always draws H (:437-440)
    else:
        H0 = bridge_h_init(seed_int, k, n)

    # ---- config scalars (build_config_from_params, src/RcppFunctions_nmf.cpp:74-76) ----
    cd_maxit = int(p["cd_maxit"]) if int(p["cd_maxit"]) > 0 else 10
    if p["cd_tol"] > 0 and p["cd_tol"] != 1e-8:
        warnings.warn("cd_tol is not carried by the GPU bridge (gpu/bridge_nmf.hpp:39-75); the engine uses 1e-8",
                      stacklevel=2)
    norm_type = {"L1": 0, "L2": 1, "none": 2}[p["norm"]]

    det = gpu_detect()
    if det["status"] != 0 or det["num_gpus"] < 1:
        raise _lib.NativeLibraryError("nmf(): no CUDA device (rcppml_gpu_detect) and no CPU fallback in this library")

    misc = dict(loss_type=loss, w_init=w_init, solver_mode=solver_mode, L1=L1, L2=L2, nonneg=nonneg,
                upper_bound=upper_bound, backend="gpu", seed_int=seed_int)

    if test_fraction > 0:
        # nmf::nmf -> dispatch_cv -> bridge_nmf_cv_sparse (nmf/fit.hpp:187-203)
        if mask_pat is not None:
            _refuse("mask = <matrix> with test_fraction > 0", "not carried by the CV entry (gpu/bridge_nmf.hpp:78-99)")
        if max(upper_bound) > 0:
            warnings.warn("upper_bound is not carried by the CV bridge (gpu/bridge_nmf.hpp:78-99) and is ignored",
                          stacklevel=2)
        cv_seed = 0 if p["cv_seed"] is None else int(np.atleast_1d(p["cv_seed"])[0])   # cv_seeds.empty() ? 0u
        if p["h_init"] is not None:
            warnings.warn("h_init is ignored by the CV bridge (gpu/bridge_nmf.hpp:437-440)", stacklevel=2)
        H0cv = bridge_h_init(seed_int, k, n)
        r = bridge_nmf_cv_sparse(A.indptr, A.indices, A.data, m, n, k, W_T0, H0cv, max_iter=int(maxit), tol=tol, L1=L1,
                                 L2=L2, nonneg=nonneg, cd_maxit=cd_maxit, verbose=verbose, seed=seed_int,
                                 holdout_fraction=float(np.float32(test_fraction)), cv_seed=cv_seed,
                                 mask_zeros=mask_zeros, norm_type=norm_type, solver_mode=solver_mode)
        if r.status != 0:
            raise _lib.NativeLibraryError("GPU CV NMF bridge call failed (status != 0)")
        misc.update(train_loss=r.train_loss, test_loss=r.best_test_loss, best_iter=r.best_iter + 1, loss=r.train_loss,
                    iter=r.iterations, tol=float("nan"), converged=r.converged)
    else:
        kw = dict(max_iter=int(maxit), tol=tol, L1=L1, L2=L2, upper_bound=upper_bound, nonneg=nonneg,
                  cd_maxit=cd_maxit, verbose=verbose, seed=seed_int, patience=5, norm_type=norm_type,
                  solver_mode=solver_mode)
        if mask_pat is not None:
            kw["mask"] = (mask_pat.indptr, mask_pat.indices)
        r = bridge_nmf_sparse(A.indptr, A.indices, A.data, m, n, k, W_T0, H0, **kw)
        if r.status != 0:
            raise _lib.NativeLibraryError("GPU NMF bridge call failed (status != 0)")
        misc.update(loss=r.train_loss, iter=r.iterations, tol=r.final_tol, converged=r.converged)

    w = r.W_T.astype(np.float64)              # result_to_list casts to double (src/RcppFunctions_nmf.cpp:109-111)
    h = r.H.astype(np.float64).T.copy()
    d = r.d.astype(np.float64)
    if sort_model:
        w, d, h = sort_by_d(w, d, h)
    misc["runtime"] = time.time() - start
    return NMFModel(np.ascontiguousarray(w), d, np.ascontiguousarray(h), misc)


# -------------------------------------------------------------------------------------------------- nnls ----
def nnls(w=None, h=None, A=None, L1=(0, 0), L2=(0, 0), loss="mse", upper_bound=(0, 0), nonneg=(True, True), threads=0,
         verbose=False, **dots):
    """Mirror of the reference's `nnls()` (R/solve.R:60-360) over `rcppml_gpu_nnls_double` (c_nnls,
    src/RcppFunctions_utils.cpp:314-366; fp64 like the reference).

    Exactly one of `w` (m x k: solve A ~ w h for h, k x n) or `h` (k x n: solve for w, m x k, as the transposed problem
    `c_nnls(t(h), t(A))`) is given. Pairs are c(w, h): solving for h uses the second element, solving for w the first
    (R/solve.R:305-309, 330-334). A factor passed in the other orientation is transposed automatically (:236-247).
    `...`: cd_maxit (100), cd_tol (1e-8), warm_start. Non-MSE losses, L21, angular and targets go through a one-iteration
    nmf() in the reference (:136-184) and are refused here."""
    from . import project
    import scipy.sparse as sp
    known = dict(L21=(0, 0), angular=(0, 0), cd_maxit=100, cd_tol=1e-8, warm_start=None, target_H=None,
                 target_lambda=(0, 0))
    unknown = [k for k in dots if k not in known]
    if unknown:
        raise TypeError("Unknown parameter(s) passed to nnls(): " + ", ".join(f"'{u}'" for u in unknown)
                        + ". See ?nnls for valid parameters.")
    known.update(dots)
    if w is not None and h is not None and A is None:                 # deprecated nnls(w, A) positional form (:79-87)
        warnings.warn("nnls(w, A) 2-positional-arg form is deprecated.\nUse nnls(w = ..., A = ...) to solve for H, or "
                      "nnls(h = ..., A = ...) to solve for W.", DeprecationWarning, stacklevel=2)
        A, h = h, None
    if w is None and h is None:
        raise ValueError("Either 'w' or 'h' must be provided (not both NULL)")
    if w is not None and h is not None:
        raise ValueError("Exactly one of 'w' or 'h' must be NULL (cannot provide both)")
    if A is None:
        raise ValueError("argument \"A\" is missing, with no default")
    solve_for_h = w is not None
    rep2 = lambda v: tuple(np.repeat(np.atleast_1d(v), 2)) if np.size(v) == 1 else tuple(np.atleast_1d(v))
    L1, L2, upper_bound, nonneg = rep2(L1), rep2(L2), rep2(upper_bound), rep2(nonneg)
    if loss != "mse" or any(np.asarray(rep2(known["L21"])) != 0) or any(np.asarray(rep2(known["angular"])) != 0) or (
            known["target_H"] is not None and any(np.asarray(rep2(known["target_lambda"])) != 0)):
        _refuse("nnls(loss / L21 / angular / target_H)")
    if isinstance(A, str):
        _refuse("A = <file path>", "file input is outside this library's scope")
    if any(v < 0 for v in L1) or any(v >= 1 for v in L1):
        raise ValueError("L1 penalty must be in range [0, 1)")
    if any(v < 0 for v in L2):
        raise ValueError("L2 penalty must be >= 0")
    if any(v < 0 for v in upper_bound):
        raise ValueError("upper_bound must be >= 0")
    A = A.tocsc() if sp.issparse(A) else sp.csc_matrix(np.asarray(A, np.float64))   # zeros add nothing to w^T a_j
    F = np.asarray(w if solve_for_h else h, np.float64)
    if F.ndim != 2:
        raise ValueError("factor matrix must be a matrix")
    m, n = A.shape
    if solve_for_h:
        if F.shape[0] != m and F.shape[1] == m:
            F = F.T
        if F.shape[0] != m:
            raise ValueError(f"Incompatible dimensions: nrow(w) = {F.shape[0]} but nrow(A) = {m} and no rownames/colnames "
                             "available for matching")
    else:
        if F.shape[1] != n and F.shape[0] == n:
            F = F.T
        if F.shape[1] != n:
            raise ValueError(f"Incompatible dimensions: ncol(h) = {F.shape[1]} but ncol(A) = {n} and no rownames/colnames "
                             "available for matching")
    ws = known["warm_start"]
    kw = dict(cd_maxit=int(known["cd_maxit"]), cd_tol=float(known["cd_tol"]))
    if solve_for_h:
        if ws is not None:
            ws = np.asarray(ws, np.float64)
            if ws.shape != (F.shape[1], n):
                raise ValueError(f"warm_start dimensions ({ws.shape[0]} x {ws.shape[1]}) don't match expected output "
                                 f"dimensions ({F.shape[1]} x {n})")
        return project.nnls(F, A, L1=float(L1[1]), L2=float(L2[1]), upper_bound=float(upper_bound[1]),
                            nonneg=bool(nonneg[1]), warm_start=ws, **kw)
    if ws is not None:
        # R/solve.R:291-292 expects nrow(A) x ncol(h) (= m x n) here and hands its transpose to c_nnls, whose
        # `G * h` then has mismatched shapes: the reference's warm start is not usable when solving for w.
        _refuse("warm_start with h = ...", "not usable in the reference either (R/solve.R:291-298 vs c_nnls)")
    At = A.T.tocsc()
    At.sort_indices()
    w_T = project.nnls(np.ascontiguousarray(F.T), At, L1=float(L1[0]), L2=float(L2[0]), upper_bound=float(upper_bound[0]),
                       nonneg=bool(nonneg[0]), **kw)
    return w_T.T.copy()
