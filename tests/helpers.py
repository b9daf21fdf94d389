"""Shared test helpers: seeded sparse inputs and the parity metric."""
import numpy as np
import scipy.sparse as sp

# north_star: "W, d, H match the reference CPU path within 1e-5 relative fp32 tolerance".
# Relative = max |gpu - oracle| over the array, divided by max |oracle| (the array's scale).
RTOL = 1e-5


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    return float(np.abs(a - b).max() / (scale if scale > 0 else 1.0))


def random_csc(m, n, density, seed, *, counts=False, ragged=False):
    """Seeded m×n CSC float32 with sorted row indices (dgCMatrix layout)."""
    rng = np.random.default_rng(seed)
    A = sp.random(m, n, density=density, format="csc", random_state=rng, dtype=np.float32)
    if counts:
        A.data = np.ceil(A.data * 8).astype(np.float32)
    else:
        A.data = (A.data + 0.5).astype(np.float32)
    if ragged:                       # some empty columns / rows and one dense column
        A = A.tolil()
        A[:, 0] = 0
        A[0, :] = 0
        A[:, n // 2] = rng.random((m, 1)).astype(np.float32) + 0.1
        A = A.tocsc()
    A.sort_indices()
    A.eliminate_zeros()
    return A


def zero_pattern_equal(a, b):
    return bool(np.array_equal(np.asarray(a) == 0, np.asarray(b) == 0))
