"""Thin object wrapper over the C ABI: node handles, host-numpy calls and device-pointer (torch) calls."""
import ctypes

import numpy as np

from . import _lib

lib = _lib.load()

KIND_GRID2D, KIND_GRID3D, KIND_GRID4D, KIND_TILEABLE2D = 2, 3, 4, 5
KIND_POS2D, KIND_POS3D, KIND_POS4D = 12, 13, 14

from .trees import TREES  # noqa: E402,F401  (pure data: importable without the native library)


def device_available():
    return bool(lib.fnb200DeviceAvailable())


def last_error():
    return lib.fnb200LastError().decode()


MATH_STRICT, MATH_RELAXED, MATH_STRICT_GENERIC, MATH_RELAXED_GENERIC = 0, 1, 2, 3


def set_math_mode(mode):
    """fnb200SetMathMode: bit 0 = relaxed arithmetic, bit 1 = generic (non-table) gradient flavour; see FastNoise_B200.h."""
    lib.fnb200SetMathMode(int(mode))


def _ptr(a):
    """numpy array -> host pointer, torch tensor -> data_ptr(), int -> as is."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return int(a)


class Generator:
    """Owns one fn* node handle."""

    def __init__(self, encoded=None, handle=None):
        if handle is None:
            handle = lib.fnNewFromEncodedNodeTree(encoded.encode() if isinstance(encoded, str) else encoded, 0xFFFFFFFF)
            if not handle:
                raise ValueError("fnNewFromEncodedNodeTree rejected %r" % (encoded,))
        self.handle = handle

    @classmethod
    def from_metadata(cls, node_id):
        handle = lib.fnNewFromMetadata(node_id, 0xFFFFFFFF)
        if not handle:
            raise ValueError("bad node id %r" % (node_id,))
        return cls(handle=handle)

    def __del__(self):
        if getattr(self, "handle", None) and lib is not None:
            lib.fnDeleteNodeRef(self.handle)
            self.handle = None

    @property
    def metadata_id(self):
        return lib.fnGetMetadataID(self.handle)

    def kernel_source(self, kind):
        n = lib.fnb200GetKernelSource(self.handle, kind, None, 0)
        if n <= 0:
            raise RuntimeError(last_error())
        buf = ctypes.create_string_buffer(n + 1)
        lib.fnb200GetKernelSource(self.handle, kind, buf, n + 1)
        return buf.value.decode()

    def kernel_params(self, kind):
        """Current values of the kernel's float-parameter argument (fnb200GetKernelParams), as a float32 array."""
        n = lib.fnb200GetKernelParams(self.handle, kind, None, 0)
        if n < 0:
            raise RuntimeError(last_error())
        buf = np.zeros(max(n, 1), np.float32)
        lib.fnb200GetKernelParams(self.handle, kind, buf.ctypes.data, n)
        return buf

    def prepare(self, kind):
        rc = lib.fnb200Prepare(self.handle, kind)
        if rc != 0:
            raise RuntimeError("fnb200Prepare(%d) = %d: %s" % (kind, rc, last_error()))

    # ---- reference-shaped calls (host numpy arrays or device tensors), return (min, max)
    @staticmethod
    def _checked(mm):
        # min > max is the failure convention, but it is also what an all-NaN output leaves behind (NaNs are skipped):
        # the error text, cleared at the start of every Gen* call, tells the two apart
        if not (mm[0] <= mm[1]) and last_error():
            raise RuntimeError("generation failed: " + last_error())
        return float(mm[0]), float(mm[1])

    def grid2d(self, out, x0, y0, nx, ny, sx, sy, seed):
        mm = np.zeros(2, np.float32)
        lib.fnGenUniformGrid2D(self.handle, _ptr(out), x0, y0, nx, ny, sx, sy, seed, mm.ctypes.data)
        return self._checked(mm)

    def grid3d(self, out, x0, y0, z0, nx, ny, nz, sx, sy, sz, seed):
        mm = np.zeros(2, np.float32)
        lib.fnGenUniformGrid3D(self.handle, _ptr(out), x0, y0, z0, nx, ny, nz, sx, sy, sz, seed, mm.ctypes.data)
        return self._checked(mm)

    def grid4d(self, out, x0, y0, z0, w0, nx, ny, nz, nw, sx, sy, sz, sw, seed):
        mm = np.zeros(2, np.float32)
        lib.fnGenUniformGrid4D(self.handle, _ptr(out), x0, y0, z0, w0, nx, ny, nz, nw, sx, sy, sz, sw, seed, mm.ctypes.data)
        return self._checked(mm)

    def tileable2d(self, out, nx, ny, sx, sy, seed):
        mm = np.zeros(2, np.float32)
        lib.fnGenTileable2D(self.handle, _ptr(out), nx, ny, sx, sy, seed, mm.ctypes.data)
        return self._checked(mm)

    def positions(self, out, pos, offsets, seed):
        dims = len(pos)
        count = len(out) if isinstance(out, np.ndarray) else out.numel()
        mm = np.zeros(2, np.float32)
        p = [_ptr(a) for a in pos]
        o = list(offsets) + [0.0] * (4 - len(offsets))
        if dims == 2:
            lib.fnGenPositionArray2D(self.handle, _ptr(out), count, p[0], p[1], o[0], o[1], seed, mm.ctypes.data)
        elif dims == 3:
            lib.fnGenPositionArray3D(self.handle, _ptr(out), count, p[0], p[1], p[2], o[0], o[1], o[2], seed, mm.ctypes.data)
        else:
            lib.fnGenPositionArray4D(self.handle, _ptr(out), count, p[0], p[1], p[2], p[3], o[0], o[1], o[2], o[3], seed, mm.ctypes.data)
        return self._checked(mm)
