"""crystals-kyber_b200: B200-native batched ML-KEM behind the API of rsjahnige/CRYSTALS-Kyber.

The product is the C-ABI shared library `libmlkem_b200.so` (CUDA kernels for sm_100a + a thin host layer;
headers in include/).  This package is the Python-side mirror used by tests and bench.py: it loads the
library with ctypes and exposes the batched entry points on numpy arrays (host memory) or torch CUDA tensors
(device memory).  There is no fallback: if the library is missing, `load()` raises.
"""
from .lib import LIB_PATH, MlKemB200Error, load  # noqa: F401
from .api import MLKEM, PARAMS, sizes  # noqa: F401
from .sharding import shard_range  # noqa: F401

__all__ = ["MLKEM", "PARAMS", "sizes", "load", "LIB_PATH", "MlKemB200Error", "shard_range"]
