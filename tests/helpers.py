"""Test-side access to the oracles (checker only; nothing here is on the product path)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def load_ref(variant):
    """variant: strict_avx2 | strict_avx512 | relaxed_avx2 | relaxed_avx512 (avx2 = generic gradient flavour, avx512 = table
    flavour). Returns a bound ctypes library, or None when it is not built or this CPU cannot run it."""
    from fastnoise2_b200 import _lib
    path = os.path.join(REF_DIR, "libfn2ref_%s.so" % variant)
    if not os.path.exists(path):
        return None
    flags = cpu_flags()
    need = {"avx2", "fma"} | ({"avx512f", "avx512dq", "avx512bw", "avx512vl"} if variant.endswith("avx512") else set())
    if flags and not need <= flags:
        return None
    lib = _lib.bind(path)
    F, I, P = ctypes.c_float, ctypes.c_int, ctypes.c_void_p
    lib.fn2refTimedGridMT.restype = ctypes.c_double
    lib.fn2refTimedGridMT.argtypes = [P, P, I, I, I, I, F, F, F, F, I, I, P]
    lib.fn2refTimedPositionsMT.restype = ctypes.c_double
    lib.fn2refTimedPositionsMT.argtypes = [P, P, I, I, P, P, P, P, I, I, P]
    if hasattr(lib, "fn2refTimedTileableMT"):  # oracle builds older than the threaded tileable driver lack it
        lib.fn2refTimedTileableMT.restype = ctypes.c_double
        lib.fn2refTimedTileableMT.argtypes = [P, P, I, I, F, F, I, I, P]
    if hasattr(lib, "fn2refCanonicalEncodedNodeTree"):       # the reference's own DeserialiseNodeData + SerialiseNodeData (oracle/ref/ref_extra.cpp)
        lib.fn2refCanonicalEncodedNodeTree.restype = ctypes.c_long
        lib.fn2refCanonicalEncodedNodeTree.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_long]
    return lib


class RefNode:
    """A node handle inside one oracle library."""

    def __init__(self, lib, encoded):
        self.lib = lib
        self.handle = lib.fnNewFromEncodedNodeTree(encoded.encode(), 0xFFFFFFFF)
        if not self.handle:
            raise ValueError("oracle rejected %r" % encoded)

    def grid2d(self, x0, y0, nx, ny, sx, sy, seed):
        out = np.empty(nx * ny, np.float32)
        mm = np.zeros(2, np.float32)
        self.lib.fnGenUniformGrid2D(self.handle, out.ctypes.data, x0, y0, nx, ny, sx, sy, seed, mm.ctypes.data)
        return out, mm

    def grid3d(self, x0, y0, z0, nx, ny, nz, sx, sy, sz, seed):
        out = np.empty(nx * ny * nz, np.float32)
        mm = np.zeros(2, np.float32)
        self.lib.fnGenUniformGrid3D(self.handle, out.ctypes.data, x0, y0, z0, nx, ny, nz, sx, sy, sz, seed, mm.ctypes.data)
        return out, mm

    def grid4d(self, x0, y0, z0, w0, nx, ny, nz, nw, sx, sy, sz, sw, seed):
        out = np.empty(nx * ny * nz * nw, np.float32)
        mm = np.zeros(2, np.float32)
        self.lib.fnGenUniformGrid4D(self.handle, out.ctypes.data, x0, y0, z0, w0, nx, ny, nz, nw, sx, sy, sz, sw, seed, mm.ctypes.data)
        return out, mm

    def tileable2d(self, nx, ny, sx, sy, seed):
        out = np.empty(nx * ny, np.float32)
        mm = np.zeros(2, np.float32)
        self.lib.fnGenTileable2D(self.handle, out.ctypes.data, nx, ny, sx, sy, seed, mm.ctypes.data)
        return out, mm

    def positions(self, pos, offsets, seed):
        n = len(pos[0])
        out = np.empty(n, np.float32)
        mm = np.zeros(2, np.float32)
        p = [a.ctypes.data for a in pos]
        o = list(offsets) + [0.0] * (4 - len(offsets))
        if len(pos) == 2:
            self.lib.fnGenPositionArray2D(self.handle, out.ctypes.data, n, p[0], p[1], o[0], o[1], seed, mm.ctypes.data)
        elif len(pos) == 3:
            self.lib.fnGenPositionArray3D(self.handle, out.ctypes.data, n, p[0], p[1], p[2], o[0], o[1], o[2], seed, mm.ctypes.data)
        else:
            self.lib.fnGenPositionArray4D(self.handle, out.ctypes.data, n, p[0], p[1], p[2], p[3], o[0], o[1], o[2], o[3], seed, mm.ctypes.data)
        return out, mm


def fnv1a_words(a):
    """32-bit FNV-1a over the 32-bit words of `a` (the checksum SURVEY.md section 8c tabulates)."""
    h = 2166136261
    for w in np.ascontiguousarray(a).view(np.uint32).tolist():
        h = ((h ^ w) * 16777619) & 0xFFFFFFFF
    return h


def compare(got, want):
    """-> dict(identical fraction, max abs diff, count over 2e-5)"""
    got = np.asarray(got, np.float32)
    want = np.asarray(want, np.float32)
    same = got.view(np.uint32) == want.view(np.uint32)
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    diff[same] = 0
    return {"identical": float(same.mean()), "max_abs": float(np.nanmax(diff)) if diff.size else 0.0,
            "over_tol": int((diff > 2e-5).sum()), "n": int(got.size)}


class ProductNode:
    """Same call surface as RefNode, backed by the product library (host numpy buffers through the fn* C ABI)."""

    def __init__(self, fn, encoded):
        self.gen = fn.Generator(encoded)

    def grid2d(self, x0, y0, nx, ny, sx, sy, seed):
        out = np.empty(nx * ny, np.float32)
        return out, np.array(self.gen.grid2d(out, x0, y0, nx, ny, sx, sy, seed), np.float32)

    def grid3d(self, x0, y0, z0, nx, ny, nz, sx, sy, sz, seed):
        out = np.empty(nx * ny * nz, np.float32)
        return out, np.array(self.gen.grid3d(out, x0, y0, z0, nx, ny, nz, sx, sy, sz, seed), np.float32)

    def grid4d(self, x0, y0, z0, w0, nx, ny, nz, nw, sx, sy, sz, sw, seed):
        out = np.empty(nx * ny * nz * nw, np.float32)
        return out, np.array(self.gen.grid4d(out, x0, y0, z0, w0, nx, ny, nz, nw, sx, sy, sz, sw, seed), np.float32)

    def tileable2d(self, nx, ny, sx, sy, seed):
        out = np.empty(nx * ny, np.float32)
        return out, np.array(self.gen.tileable2d(out, nx, ny, sx, sy, seed), np.float32)

    def positions(self, pos, offsets, seed):
        out = np.empty(len(pos[0]), np.float32)
        return out, np.array(self.gen.positions(out, pos, offsets, seed), np.float32)


def load_golden(suffix=""):
    """suffix "" = generic gradient flavour, "_tables" = table flavour (tests/conftest.py FLAVOURS)."""
    import json
    golden_dir = os.path.join(ROOT, "tests", "golden")
    outputs = np.load(os.path.join(golden_dir, "outputs%s.npz" % suffix))
    summary = json.load(open(os.path.join(golden_dir, "summary%s.json" % suffix)))
    metadata = json.load(open(os.path.join(golden_dir, "metadata.json")))
    return outputs, summary, metadata


def load_port():
    """oracle/libfn2oracle.so: the plain-C restatement (oracle/fn2_oracle.c). None if not built."""
    path = os.path.join(ROOT, "oracle", "libfn2oracle.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    P, I, F, LL = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
    lib.fn2oDecode.restype = P
    lib.fn2oDecode.argtypes = [ctypes.c_char_p]
    lib.fn2oFree.argtypes = [P]
    lib.fn2oSupported.argtypes = [P]
    lib.fn2oSetConvertMode.argtypes = [I]
    lib.fn2oNodeId.argtypes = [P]
    lib.fn2oGrid.argtypes = [P, P, I, P, P, P, I, LL, LL, P]
    lib.fn2oTileable.argtypes = [P, P, I, I, F, F, I, I, I, P]
    lib.fn2oPositions.argtypes = [P, P, I, LL, P, P, P, P, P, I, P]
    if hasattr(lib, "fn2oDebugGrid3D"):
        lib.fn2oDebugGrid3D.argtypes = [P, P, P, P, P, I]
    return lib


class PortNode:
    """Same call surface as RefNode over the C restatement; `rows=(begin, count)` evaluates a slab with global indices."""

    def __init__(self, lib, encoded):
        self.lib = lib
        self.handle = lib.fn2oDecode(encoded.encode())
        self.supported = bool(self.handle) and bool(lib.fn2oSupported(self.handle))

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.fn2oFree(self.handle)

    def _grid(self, dims, off, n, step, seed, rows=None):
        total_rows = int(np.prod(n[1:])) if dims > 1 else 1
        begin, count = rows if rows else (0, total_rows)
        out = np.empty(count * n[0], np.float32)
        mm = np.zeros(2, np.float32)
        offs = np.array(list(off) + [0.0] * (4 - dims), np.float32)
        ns = np.array(list(n) + [1] * (4 - dims), np.int32)
        steps = np.array(list(step) + [0.0] * (4 - dims), np.float32)
        self.lib.fn2oGrid(self.handle, out.ctypes.data, dims, offs.ctypes.data, ns.ctypes.data, steps.ctypes.data, seed, begin, count, mm.ctypes.data)
        return out, mm

    def grid2d(self, x0, y0, nx, ny, sx, sy, seed, rows=None):
        return self._grid(2, (x0, y0), (nx, ny), (sx, sy), seed, rows)

    def grid3d(self, x0, y0, z0, nx, ny, nz, sx, sy, sz, seed, rows=None):
        return self._grid(3, (x0, y0, z0), (nx, ny, nz), (sx, sy, sz), seed, rows)

    def grid4d(self, x0, y0, z0, w0, nx, ny, nz, nw, sx, sy, sz, sw, seed, rows=None):
        return self._grid(4, (x0, y0, z0, w0), (nx, ny, nz, nw), (sx, sy, sz, sw), seed, rows)

    def tileable2d(self, nx, ny, sx, sy, seed, rows=None):
        begin, count = rows if rows else (0, ny)
        out = np.empty(count * nx, np.float32)
        mm = np.zeros(2, np.float32)
        self.lib.fn2oTileable(self.handle, out.ctypes.data, nx, ny, sx, sy, seed, begin, count, mm.ctypes.data)
        return out, mm

    def positions(self, pos, offsets, seed):
        n = len(pos[0])
        out = np.empty(n, np.float32)
        mm = np.zeros(2, np.float32)
        p = [a.ctypes.data for a in pos] + [None] * (4 - len(pos))
        offs = np.array(list(offsets) + [0.0] * (4 - len(offsets)), np.float32)
        self.lib.fn2oPositions(self.handle, out.ctypes.data, len(pos), n, p[0], p[1], p[2], p[3], offs.ctypes.data, seed, mm.ctypes.data)
        return out, mm
