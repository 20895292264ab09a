"""cortex_b200 -- B200-native similarity scan behind cortex-core's VectorIndex surface.

Only what the hot path needs lives here: csrc/ (CUDA kernels + C ABI), the ctypes
declarations of that ABI, and the host-side mirror of the reference interface.
"""
from .index import CortexError, GpuVectorIndex, SimilarityResult, VectorFilter  # noqa: F401
from .config import SimilarityConfig  # noqa: F401

__all__ = ["CortexError", "GpuVectorIndex", "SimilarityResult", "VectorFilter", "SimilarityConfig"]
