"""fastnoise2_b200 -- B200-native (sm_100a) evaluation of FastNoise2 node trees.

Python here is plumbing for tests, the benchmark and torch.distributed; the product is
libfastnoise2_b200.so (C++ host mirror of the FastNoise API + NVRTC-specialised CUDA kernels), reached
through the C ABI declared in include/FastNoise/FastNoise_C.h and FastNoise_B200.h.
"""
from . import _lib
from .api import Generator, lib, TREES, device_available, last_error, set_math_mode  # noqa: F401
from .api import MATH_STRICT, MATH_RELAXED, MATH_STRICT_GENERIC, MATH_RELAXED_GENERIC  # noqa: F401
