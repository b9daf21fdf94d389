"""rcppml_b200 — B200-native (sm_100a) sparse-NMF ALS engine behind RcppML's GPU bridge.

Only the ALS hot path of the reference lives here: the CUDA engine in csrc/
(built to lib/RcppML_gpu.so, exporting the reference's C ABI) and the host-side mirror of the
reference interface for this path.
"""
from .engine import Engine, FitResult, make_config, nccl_unique_id  # noqa: F401
from .bridge import bridge_nmf_cv_sparse, bridge_nmf_sparse, gpu_detect, gpu_nmf_zerocopy  # noqa: F401
from .nmf import NMFModel, nmf, nnls  # noqa: F401
from .project import evaluate, predict  # noqa: F401
from .streampress import (GpuSparseMatrix, SpzFile, nmf_zerocopy, st_free_gpu, st_info, st_read,  # noqa: F401
                          st_read_gpu, st_read_transpose)

__all__ = ["Engine", "FitResult", "make_config", "nccl_unique_id", "bridge_nmf_sparse", "bridge_nmf_cv_sparse", "gpu_detect", "gpu_nmf_zerocopy", "nmf", "NMFModel", "nnls", "predict", "evaluate",
           "SpzFile", "GpuSparseMatrix", "st_info", "st_read", "st_read_transpose", "st_read_gpu", "st_free_gpu", "nmf_zerocopy"]
