"""fastnoise2_b200 -- B200-native (sm_100a) evaluation of FastNoise2 node trees.

Python here is plumbing for tests, the benchmark and torch.distributed; the product is
libfastnoise2_b200.so (C++ host mirror of the FastNoise API + NVRTC-specialised CUDA kernels), reached
through the C ABI declared in include/FastNoise/FastNoise_C.h and FastNoise_B200.h.

The native library is loaded on first use of anything that needs it (`fastnoise2_b200.lib`, `Generator`, ...);
`fastnoise2_b200.trees` and `fastnoise2_b200._lib` (prototype table, used to bind the oracle libraries) are pure Python.
"""
from .trees import TREES  # noqa: F401

_API = ("Generator", "lib", "device_available", "last_error", "set_math_mode",
        "MATH_STRICT", "MATH_RELAXED", "MATH_STRICT_GENERIC", "MATH_RELAXED_GENERIC")


def __getattr__(name):
    if name in _API or name == "api":
        import importlib
        api = importlib.import_module(__name__ + ".api")
        return api if name == "api" else getattr(api, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
