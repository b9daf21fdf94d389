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


def load_pbmc3k():
    """The reference's real dataset (inst/extdata/pbmc3k.spz, 13714 x 2700 counts), decoded by the REFERENCE's
    own StreamPress reader at build time into oracle/_ref/pbmc3k.bin (`make -C oracle ref`; git-ignored but it
    travels to the GPU box). Returns a scipy CSC matrix or None when the fixture is not present."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "pbmc3k.bin")
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        m, n = np.fromfile(f, np.int32, 2)
        nnz = int(np.fromfile(f, np.int64, 1)[0])
        p = np.fromfile(f, np.int32, n + 1)
        i = np.fromfile(f, np.int32, nnz)
        x = np.fromfile(f, np.float32, nnz)
    A = sp.csc_matrix((x, i, p), shape=(int(m), int(n)))
    A.sort_indices()
    return A


def load_pbmc3k_block():
    """tests/golden/pbmc3k_500x200.npz: rows 0..499, columns 0..199 of pbmc3k — the block the reference's own
    GPU accuracy test uses (tests/testthat/test_gpu_accuracy.R:26-34). Committed; made by
    tests/golden/make_golden.py from oracle/_ref/pbmc3k.bin."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pbmc3k_500x200.npz"))
    return sp.csc_matrix((g["data"], g["indices"], g["indptr"]), shape=(500, 200))


def load_movielens():
    """tests/golden/movielens.npz: the reference's data/movielens.rda (Matrix::dgCMatrix, 3867 x 610, 75 238 ratings
    1..5) read without R by tests/golden/rdx3.py — BASELINE.json configs[1]. Returns (csc, frozen oracle outputs)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "movielens.npz"))
    A = sp.csc_matrix((g["data"], g["indices"], g["indptr"]), shape=(int(g["m"]), int(g["n"])))
    return A, g


def load_aml_as_csc():
    """tests/golden/aml.npz: the reference's data/aml.rda (dense 824 x 135 methylation matrix, BASELINE.json
    configs[0], the quick-start) as CSC with its 232 exact zeros dropped — the sparse path's view of it."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aml.npz"))
    A = sp.csc_matrix(g["A"])
    A.sort_indices()
    return A
