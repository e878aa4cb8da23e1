"""Parity at the sizes BASELINE.json states, word for word against the strict oracle, plus a direct comparison of
the INTEGER intermediates (primed cell coordinates, HashPrimesHB, Cellular winner) with the C restatement's.

The oracle for the full-size runs is oracle/_ref strict (the reference's own node sources) driven over all host cores in chunks
of the outermost axis, each chunk through the ordinary fnGenUniformGrid* call with a FLOAT chunk offset; the product generates
the same chunk through the slab entry point with GLOBAL integer indices. The two agree bit for bit only because every position
of these configs is exactly representable (offset 0, step 1 or 0.25: fma( float(i), step, 0 ) == float(i - k) * step + k * step);
the test asserts that precondition. Each config prints its mismatch count; the bar is 0.
"""
import ctypes
import json
import os
import time

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu
SEED = 1337


@pytest.fixture(scope="module")
def gpu(fn):
    if not fn.device_available():
        pytest.skip("no CUDA device")
    return fn


def strict_oracle(fn):
    """(library, product math mode): the table flavour when this CPU can run the AVX-512 build of the reference, else the generic one."""
    lib = helpers.load_ref("strict_avx512")
    if lib is not None:
        return lib, 0, "tables"
    lib = helpers.load_ref("strict_avx2")
    if lib is None:
        pytest.skip("oracle/_ref is not built")
    return lib, 2, "generic"


def report(name, total, mismatches, seconds, extra=""):
    line = "full-size parity %-28s %12d points  %d mismatches  (%.1f s)%s" % (name, total, mismatches, seconds, extra)
    print("\n" + line, flush=True)
    out_dir = os.path.join(helpers.ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "fullsize_parity.txt"), "a") as f:
            f.write(line + "\n")


GRID_CASES = {
    # name: (tree, dims, n, step, chunk of the outermost axis)
    "c1_simplex_256^3": ("c1_simplex", 3, (256, 256, 256), 1.0, 256),
    "c2a_fbm5_perlin_16384^2": ("c2_fbm5_perlin", 2, (16384, 16384), 0.25, 2048),
    "c3a_cellular_distance_512^3": ("c3a_cellular_distance", 3, (512, 512, 512), 1.0, 128),
    "c3b_cellular_lookup_white_512^3": ("c3b_cellular_lookup_white", 3, (512, 512, 512), 1.0, 128),
    "c3c_cellular_lookup_gradient_512^3": ("c3c_cellular_lookup_gradient", 3, (512, 512, 512), 1.0, 128),
    "c4_terrain_1024^3": ("c4_terrain", 3, (1024, 1024, 1024), 1.0, 64),
}


@pytest.mark.parametrize("name", sorted(GRID_CASES))
def test_uniform_grid_configs_at_baseline_size(gpu, name):
    import torch
    tree, dims, n, step, chunk = GRID_CASES[name]
    ref, mode, flavour = strict_oracle(gpu)
    gpu.set_math_mode(mode)
    try:
        # precondition of comparing a float-offset chunk with a global-index slab: every index * step is exact in float
        assert float(np.float32(max(n) * step)) == max(n) * step and max(n) < (1 << 22)
        g = gpu.Generator(gpu.TREES[tree])
        node = helpers.RefNode(ref, gpu.TREES[tree])
        threads = os.cpu_count() or 1
        per_outer = int(np.prod(n[:-1]))
        outer = n[-1]
        dev = torch.empty(per_outer * chunk, dtype=torch.float32, device="cuda")
        acc = torch.zeros(2, dtype=torch.int32, device="cuda")
        want = np.empty(per_outer * chunk, np.float32)
        wmm = np.zeros(2, np.float32)
        lo, hi = np.inf, -np.inf
        mismatches = 0
        t0 = time.time()
        for begin in range(0, outer, chunk):
            end = min(outer, begin + chunk)
            count = per_outer * (end - begin)
            if dims == 3:
                rc = gpu.lib.fnb200GenUniformGrid3D(g.handle, dev.data_ptr(), 0.0, 0.0, 0.0, n[0], n[1], n[2], step, step, step, SEED, begin, end, acc.data_ptr(), None)
                ref.fn2refTimedGridMT(node.handle, want.ctypes.data, 3, n[0], n[1], end - begin, 0.0, 0.0, begin * step, step, SEED, threads, wmm.ctypes.data)
            else:
                rc = gpu.lib.fnb200GenUniformGrid2D(g.handle, dev.data_ptr(), 0.0, 0.0, n[0], n[1], step, step, SEED, begin, end, acc.data_ptr(), None)
                ref.fn2refTimedGridMT(node.handle, want.ctypes.data, 2, n[0], end - begin, 1, 0.0, begin * step, 0.0, step, SEED, threads, wmm.ctypes.data)
            assert rc == 0, gpu.last_error()
            got = dev[:count].cpu().numpy()
            mismatches += int(np.count_nonzero(got.view(np.uint32) != want[:count].view(np.uint32)))
            lo, hi = min(lo, float(wmm[0])), max(hi, float(wmm[1]))
        mm = np.zeros(2, np.float32)
        assert gpu.lib.fnb200MinMaxRead(acc.data_ptr(), None, mm.ctypes.data) == 0
        report(name + " [" + flavour + "]", per_outer * outer, mismatches, time.time() - t0)
        assert mismatches == 0
        assert (float(mm[0]), float(mm[1])) == (lo, hi)
    finally:
        gpu.set_math_mode(0)


def test_tileable_config_at_baseline_size(gpu):
    """C2b: FBm5(Perlin) GenTileable2D 4096^2. The torus mapping goes through FS::Cos / FS::Sin, which exist only as this repo's
    restatement of the absent FastSIMD (oracle/shim): bit-exact against the oracle, parity unpinned upstream for this call."""
    import torch
    ref, mode, flavour = strict_oracle(gpu)
    gpu.set_math_mode(mode)
    try:
        n = 4096
        g = gpu.Generator(gpu.TREES["c2_fbm5_perlin"])
        t0 = time.time()
        dev = torch.empty(n * n, dtype=torch.float32, device="cuda")
        mm = g.tileable2d(dev, n, n, 1.0, 1.0, SEED)
        want, wmm = helpers.RefNode(ref, gpu.TREES["c2_fbm5_perlin"]).tileable2d(n, n, 1.0, 1.0, SEED)
        mismatches = int(np.count_nonzero(dev.cpu().numpy().view(np.uint32) != want.view(np.uint32)))
        report("c2b_fbm5_perlin_tileable_4096^2 [" + flavour + "]", n * n, mismatches, time.time() - t0)
        assert mismatches == 0 and mm == (float(wmm[0]), float(wmm[1]))
    finally:
        gpu.set_math_mode(0)


def test_position_array_config_at_baseline_size(gpu):
    """C5: SuperSimplex GenPositionArray4D on all 256 Mi Philox positions in U[-1000, 1000)^4, 32 Mi at a time."""
    import torch
    ref, mode, flavour = strict_oracle(gpu)
    gpu.set_math_mode(mode)
    try:
        total, chunk = 256 * 1024 * 1024, 32 * 1024 * 1024
        g = gpu.Generator(gpu.TREES["c5_supersimplex"])
        node = helpers.RefNode(ref, gpu.TREES["c5_supersimplex"])
        threads = os.cpu_count() or 1
        pos = [torch.empty(chunk, dtype=torch.float32, device="cuda") for _ in range(4)]
        out = torch.empty(chunk, dtype=torch.float32, device="cuda")
        acc = torch.zeros(2, dtype=torch.int32, device="cuda")
        want = np.empty(chunk, np.float32)
        wmm = np.zeros(2, np.float32)
        lo, hi, mismatches = np.inf, -np.inf, 0
        t0 = time.time()
        for first in range(0, total, chunk):
            for axis in range(4):
                assert gpu.lib.fnb200FillUniformPhilox(pos[axis].data_ptr(), chunk, first, SEED, axis, -1000.0, 1000.0, None) == 0
            rc = gpu.lib.fnb200GenPositionArray(g.handle, 4, out.data_ptr(), chunk, pos[0].data_ptr(), pos[1].data_ptr(), pos[2].data_ptr(), pos[3].data_ptr(),
                                                0.0, 0.0, 0.0, 0.0, SEED, acc.data_ptr(), None)
            assert rc == 0, gpu.last_error()
            host = [p.cpu().numpy() for p in pos]
            assert first > 0 or (min(h.min() for h in host) >= -1000.0 and max(h.max() for h in host) < 1000.0)
            ref.fn2refTimedPositionsMT(node.handle, want.ctypes.data, 4, chunk, host[0].ctypes.data, host[1].ctypes.data, host[2].ctypes.data, host[3].ctypes.data,
                                       SEED, threads, wmm.ctypes.data)
            mismatches += int(np.count_nonzero(out.cpu().numpy().view(np.uint32) != want.view(np.uint32)))
            lo, hi = min(lo, float(wmm[0])), max(hi, float(wmm[1]))
        mm = np.zeros(2, np.float32)
        assert gpu.lib.fnb200MinMaxRead(acc.data_ptr(), None, mm.ctypes.data) == 0
        report("c5_supersimplex_4d_256Mi_positions [" + flavour + "]", total, mismatches, time.time() - t0)
        assert mismatches == 0 and (float(mm[0]), float(mm[1])) == (lo, hi)
    finally:
        gpu.set_math_mode(0)


@pytest.mark.parametrize("case", ["c1_simplex", "c3_cellular", "c3_cellular_far"])
def test_integer_intermediates_match_the_oracle(gpu, case):
    """Hash / primed-cell / Cellular-winner intermediates, compared directly: the GPU writes them through the debug sink of the very
    device functions the generated kernels call (fnb200DebugIntermediates), the C restatement through its own (fn2oDebugGrid3D)."""
    port = helpers.load_port()
    if port is None or not hasattr(port, "fn2oDebugGrid3D"):
        pytest.skip("oracle/libfn2oracle.so not built")
    n = (96, 64, 48)
    off = (-37.25, 11.5, 3.0) if case != "c3_cellular_far" else (1.0e6, -2.5e6, 7.7e5)
    step = 17.0 if case != "c1_simplex" else 12.75  # frequency 0.01: the grid spans ~16 x 11 x 8 cells
    tree = gpu.TREES["c1_simplex"] if case == "c1_simplex" else gpu.TREES["c3b_cellular_lookup_white"]
    mode = 0 if case == "c1_simplex" else 1
    total = n[0] * n[1] * n[2]
    frequency = float(np.float32(1.0) / np.float32(100.0))  # Feature Scale 100 (the default of both nodes) -> frequency, in float
    got = np.zeros(total * 8, np.int32)
    rc = gpu.lib.fnb200DebugIntermediates(mode, SEED, frequency, 1.0, n[0], n[1], n[2], off[0], off[1], off[2], step, got.ctypes.data)
    assert rc == 0, gpu.last_error()
    node = helpers.PortNode(port, tree)
    want = np.zeros(total * 8, np.int32)
    port.fn2oDebugGrid3D(node.handle, want.ctypes.data, (ctypes.c_float * 3)(*off), (ctypes.c_int * 3)(*n), (ctypes.c_float * 3)(step, step, step), SEED)
    got, want = got.reshape(total, 8), want.reshape(total, 8)
    assert len(np.unique(want[:, 3])) > 8  # the grid crosses many cells
    bad = int(np.count_nonzero((got[:, :7] != want[:, :7]).any(axis=1)))
    if mode == 0:  # the product the gradient table lookup consumes is vertex 0's HashPrimesHB times the table stride (mod 2^32)
        assert np.array_equal(got[:, 7].view(np.uint32), (want[:, 3].view(np.uint32).astype(np.uint64) * 8 % (1 << 32)).astype(np.uint32))
    report("integer intermediates " + case, total, bad, 0.0, "  (8 ints per point: primed cell, HashPrimesHB, winner)")
    assert bad == 0
