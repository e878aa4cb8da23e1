"""Parity of the CUDA path against the oracle, through the C ABI (needs a B200: pytest -m gpu).

Bar: every output word bit-identical to the strict oracle (same operation sequence, FMAs only where the reference
has them, IEEE div/sqrt) -- which subsumes "integer hashes and cell selection bit-exact" -- with a float tolerance of
2e-5 absolute stated for the relaxed math mode (FMA contraction, measured envelope 9.1e-6: SURVEY.md 8d).
Every test that takes `ref` or `flavour` runs twice: table gradients (product default) against the reference's AVX512_F
code path, and generic gradients against its select/xor code path (tests/conftest.py FLAVOURS).
"""
import os

import numpy as np
import pytest

from tests import cases, helpers, trees

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture(scope="module")
def gpu(fn):
    assert fn.device_available(), "these tests need a CUDA device: " + fn.last_error()
    return fn


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_golden_vectors_bit_exact(gpu, flavour):
    outputs, summary, _ = helpers.load_golden(flavour.golden)
    before = gpu.lib.fnb200LaunchCount()
    for name, case in cases.GOLDEN_CASES.items():
        out, mm = cases.run_case(helpers.ProductNode(gpu, case["tree"]), case)
        assert np.array_equal(bits(out), bits(outputs[name])), (name, helpers.compare(out, outputs[name]))
        assert float(mm[0]) == summary[name]["min"] and float(mm[1]) == summary[name]["max"], name
    assert gpu.lib.fnb200LaunchCount() - before >= len(cases.GOLDEN_CASES)  # the CUDA path ran, nothing else can


@pytest.mark.parametrize("name", ["c1_simplex", "c2_fbm5_perlin", "c3a_cellular_distance", "c3b_cellular_lookup_white",
                                  "c3c_cellular_lookup_gradient", "c4_terrain", "c5_supersimplex", "fbm_simplex_readme",
                                  "mountain_terrain", "cellular_caves"])
def test_config_trees_vs_strict_oracle(gpu, ref, name):
    """Seeded grids at sizes the oracle finishes in seconds; the config's own step plus a coarse one that crosses many cells."""
    tree = {**gpu.TREES, **cases.DEMO_TREES}[name]
    shapes = {
        "c2_fbm5_perlin": [dict(kind="grid2d", off=(0.0, 0.0), n=(1024, 512), step=(0.25, 0.25)), dict(kind="grid2d", off=(-333.0, 90.5), n=(700, 300), step=(3.3, 3.3)),
                           dict(kind="tileable2d", n=(256, 192), step=(1.0, 1.0))],
        "fbm_simplex_readme": [dict(kind="grid2d", off=(0.0, 0.0), n=(512, 512), step=(1.0, 1.0))],
        "c5_supersimplex": [dict(kind="pos4d", count=300007, off=(0.0, 0.0, 0.0, 0.0)), dict(kind="grid4d", off=(0.0, 0.0, 0.0, 0.0), n=(32, 32, 16, 16), step=(4.0,) * 4)],
    }.get(name, [dict(kind="grid3d", off=(0.0, 0.0, 0.0), n=(128, 128, 64), step=(1.0, 1.0, 1.0)),
                 dict(kind="grid3d", off=(-250.5, 77.25, 1e3), n=(100, 60, 50), step=(7.9, 7.9, 7.9))])
    for shape in shapes:
        for seed in (1337, -1):
            case = dict(shape, tree=tree, seed=seed)
            got, gmm = cases.run_case(helpers.ProductNode(gpu, tree), case)
            want, wmm = cases.run_case(helpers.RefNode(ref, tree), case)
            assert np.array_equal(bits(got), bits(want)), (name, shape, seed, helpers.compare(got, want))
            assert float(gmm[0]) == float(wmm[0]) and float(gmm[1]) == float(wmm[1])


def test_every_node_type_vs_strict_oracle(gpu, ref):
    """All node types this build lowers, auto-wired through reflection, 2D/3D/4D, hybrids as constants and as generators."""
    def gen(lib, h, dim, n, seed):
        out = np.empty(n ** dim, np.float32)
        mm = np.zeros(2, np.float32)
        if dim == 2:
            lib.fnGenUniformGrid2D(h, out.ctypes.data, -3.3, 2.1, n, n, 1.9, 1.9, seed, mm.ctypes.data)
        elif dim == 3:
            lib.fnGenUniformGrid3D(h, out.ctypes.data, -3.3, 2.1, 0.7, n, n, n, 1.9, 1.9, 1.9, seed, mm.ctypes.data)
        else:
            lib.fnGenUniformGrid4D(h, out.ctypes.data, -3.3, 2.1, 0.7, -1.9, n, n, n, n, 1.9, 1.9, 1.9, 1.9, seed, mm.ctypes.data)
        return out

    failures = []
    for hybrid_nodes in (False, True):
        for node_id in range(47):
            if node_id in trees.UNSUPPORTED_IDS:
                continue
            for dim, n in ((2, 64), (3, 20), (4, 9)):
                a, b = trees.Tree(gpu.lib), trees.Tree(ref)
                got = gen(gpu.lib, a.wire(node_id, hybrid_nodes), dim, n, 1337)
                want = gen(ref, b.wire(node_id, hybrid_nodes), dim, n, 1337)
                if not np.array_equal(bits(got), bits(want)) and not (np.isnan(got).all() and np.isnan(want).all()):
                    failures.append((node_id, dim, hybrid_nodes, helpers.compare(got, want)))
                a.close()
                b.close()
    assert not failures, failures


def test_parameter_variants_vs_strict_oracle(gpu, ref):
    """Uniform branches of the reference that the generated program resolves at lowering time."""
    variants = []
    for df in range(6):                    # every distance function, Cellular nodes + DistanceToPoint
        for node_id in (10, 11, 12, 5):
            variants.append((node_id, [("Distance Function", df)], []))
    for ret in range(5):                   # CellularDistance return types with indices 1,3
        variants.append((11, [("Return Type", ret), ("Distance Index 0", 1), ("Distance Index 1", 3)], []))
    variants.append((11, [("Distance Function", 0), ("Distance Index 0", 2), ("Distance Index 1", 2)], []))
    for idx in range(4):
        variants.append((10, [("Value Index", idx)], []))
    for interp in range(3):
        variants.append((40, [("Interpolation", interp)], []))
    for dim_removed in range(4):
        variants.append((44, [("Remove Dimension", dim_removed)], []))
    variants.append((46, [("Rotation Type", 1)], []))
    for warp_id in (16, 17):               # DomainWarpSimplex / SuperSimplex: both vectorization schemes, scaled axes
        for scheme in (0, 1):
            variants.append((warp_id, [("Vectorization Scheme", scheme), ("Seed Offset", 11)], [("Feature Scale", 31.0)]))
    variants.append((41, [("Clamp Output", 1)], []))
    variants.append((32, [("Pow", 5)], []))
    variants.append((13, [("Octaves", 7)], [("Lacunarity", 2.37)]))
    variants.append((15, [("Octaves", 6)], [("Lacunarity", 1.83)]))
    variants.append((19, [("Octaves", 4)], []))
    variants.append((20, [("Octaves", 4)], []))
    variants.append((35, [], [("Yaw", 0.7)]))
    variants.append((35, [], [("Yaw", 0.7), ("Pitch", -1.1), ("Roll", 2.9)]))
    variants.append((42, [], [("Step Count", 7.0)]))
    variants.append((6, [("Seed Offset", -12345)], [("Output Min", 3.0), ("Output Max", -2.0), ("Feature Scale", 0.37)]))

    def build(lib, node_id, ints, floats, hybrid_values):
        t = trees.Tree(lib)
        h = t.wire(node_id, False)
        for name, v in ints:
            t.set_int(h, name, v)
        for name, v in floats:
            t.set_float(h, name, v)
        for i, v in hybrid_values:
            assert lib.fnSetHybridFloat(h, i, v)
        return t, h

    failures = []
    for node_id, ints, floats in variants:
        hybrid_values = []
        if node_id in (10, 11, 12):
            hybrid_values = [(1, 0.83), (2, 0.35)]  # grid jitter, size jitter active
        if node_id == 42:
            hybrid_values = [(0, 0.3)]
        if node_id in (13, 15, 19, 20):
            hybrid_values = [(0, 0.61), (1, 0.45)]   # gain, weighted strength active
        for dim, n in ((2, 48), (3, 16), (4, 8)):
            (ta, ha), (tb, hb) = build(gpu.lib, node_id, ints, floats, hybrid_values), build(ref, node_id, ints, floats, hybrid_values)
            outs = []
            for lib, h in ((gpu.lib, ha), (ref, hb)):
                out = np.empty(n ** dim, np.float32)
                if dim == 2:
                    lib.fnGenUniformGrid2D(h, out.ctypes.data, 0.4, -7.7, n, n, 2.3, 2.3, 99, None)
                elif dim == 3:
                    lib.fnGenUniformGrid3D(h, out.ctypes.data, 0.4, -7.7, 5.5, n, n, n, 2.3, 2.3, 2.3, 99, None)
                else:
                    lib.fnGenUniformGrid4D(h, out.ctypes.data, 0.4, -7.7, 5.5, 1.25, n, n, n, n, 2.3, 2.3, 2.3, 2.3, 99, None)
                outs.append(out)
            if not np.array_equal(bits(outs[0]), bits(outs[1])):
                failures.append((node_id, ints, floats, dim, helpers.compare(outs[0], outs[1])))
            ta.close()
            tb.close()
    assert not failures, failures


def test_ragged_shapes_tails_and_seeds(gpu, ref):
    """The cases the reference's AxisReset / LoadRemaining / StoreRemaining code exists for (Generator.inl:447-510)."""
    tree = gpu.TREES["c1_simplex"]
    a, b = helpers.ProductNode(gpu, tree), helpers.RefNode(ref, tree)
    for n in [(1, 1, 1), (2, 1, 1), (3, 5, 2), (7, 3, 1), (1, 9, 4), (5, 1, 7), (33, 2, 2), (4, 4, 4), (130, 3, 1)]:
        for seed in (0, 1, -1, 0x7fffffff):
            got, gmm = a.grid3d(0.5, -0.25, 3.0, *n, 11.0, 13.0, 17.0, seed)
            want, wmm = b.grid3d(0.5, -0.25, 3.0, *n, 11.0, 13.0, 17.0, seed)
            assert np.array_equal(bits(got), bits(want)), (n, seed)
            assert tuple(gmm) == tuple(wmm)
    for n in [(1, 1), (3, 1), (1, 5), (6, 7), (17, 2)]:
        got, _ = a.grid2d(0.0, 0.0, *n, 9.0, 9.0, 5)
        want, _ = b.grid2d(0.0, 0.0, *n, 9.0, 9.0, 5)
        assert np.array_equal(bits(got), bits(want)), n
    for n in [(1, 1, 1, 1), (3, 2, 2, 2), (5, 1, 3, 2)]:
        got, _ = a.grid4d(0.0, 0.0, 0.0, 0.0, *n, 9.0, 9.0, 9.0, 9.0, 5)
        want, _ = b.grid4d(0.0, 0.0, 0.0, 0.0, *n, 9.0, 9.0, 9.0, 9.0, 5)
        assert np.array_equal(bits(got), bits(want)), n
    rng = np.random.default_rng(3)
    for count in list(range(1, 20)) + [31, 32, 33, 63, 64, 65, 1000, 1001, 1002, 1003]:
        for dims in (2, 3, 4):
            pos = [rng.uniform(-500, 500, count).astype(np.float32) for _ in range(dims)]
            off = [1.5, -2.5, 3.5, -4.5][:dims]
            got, gmm = a.positions(pos, off, 1337)
            want, wmm = b.positions(pos, off, 1337)
            assert np.array_equal(bits(got), bits(want)), (count, dims)
            assert tuple(gmm) == tuple(wmm)
    t2 = helpers.ProductNode(gpu, gpu.TREES["c2_fbm5_perlin"])
    r2 = helpers.RefNode(ref, gpu.TREES["c2_fbm5_perlin"])
    for n in [(1, 1), (2, 3), (5, 4), (31, 7)]:
        got, _ = t2.tileable2d(*n, 1.5, 0.5, 1337)
        want, _ = r2.tileable2d(*n, 1.5, 0.5, 1337)
        assert np.array_equal(bits(got), bits(want)), n


def test_single_point_calls(gpu, ref):
    g, r = gpu.Generator(gpu.TREES["c4_terrain"]), helpers.RefNode(ref, gpu.TREES["c4_terrain"])
    for p in [(0.0, 0.0, 0.0), (12.5, -7.25, 3.0), (1e4, -1e4, 0.125)]:
        assert gpu.lib.fnGenSingle3D(g.handle, *p, 1337) == ref.fnGenSingle3D(r.handle, *p, 1337)
    assert gpu.lib.fnGenSingle2D(g.handle, 5.0, 6.0, 1) == ref.fnGenSingle2D(r.handle, 5.0, 6.0, 1)
    assert gpu.lib.fnGenSingle4D(g.handle, 5.0, 6.0, 7.0, 8.0, 1) == ref.fnGenSingle4D(r.handle, 5.0, 6.0, 7.0, 8.0, 1)


def test_non_finite_positions_do_not_trap(gpu):
    g = gpu.Generator(gpu.TREES["c4_terrain"])
    pos = [np.array([np.nan, np.inf, -np.inf, 1e30, -1e30, 0.0, 3e9, -3e9], np.float32) for _ in range(3)]
    out = np.zeros(8, np.float32)
    mm = np.zeros(2, np.float32)
    gpu.lib.fnGenPositionArray3D(g.handle, out.ctypes.data, 8, pos[0].ctypes.data, pos[1].ctypes.data, pos[2].ctypes.data, 0, 0, 0, 1337, mm.ctypes.data)
    assert np.isfinite(out[5])  # and the call returned


def test_host_staging_matches_device_buffers(gpu):
    """Chunked host pipeline (tiny staging buffers force many chunks) == one in-place launch on a device buffer."""
    import subprocess
    import sys
    code = r'''
import os, sys
os.environ["FASTNOISE2_B200_STAGING_MB"] = "1"
sys.path.insert(0, %r)
import numpy as np, torch
import fastnoise2_b200 as fn
g = fn.Generator(fn.TREES["c4_terrain"])
n = (200, 75, 40)
host = np.empty(n[0] * n[1] * n[2], np.float32)
mm_h = g.grid3d(host, 0, 0, 0, *n, 1.3, 1.3, 1.3, 1337)
dev = torch.empty(host.size, dtype=torch.float32, device="cuda")
mm_d = g.grid3d(dev, 0, 0, 0, *n, 1.3, 1.3, 1.3, 1337)
assert mm_h == mm_d, (mm_h, mm_d)
assert np.array_equal(host.view(np.uint32), dev.cpu().numpy().view(np.uint32))
s = fn.Generator(fn.TREES["c5_supersimplex"])
rng = np.random.default_rng(0)
cnt = 700001
pos = [rng.uniform(-1000, 1000, cnt).astype(np.float32) for _ in range(4)]
out_h = np.empty(cnt, np.float32)
mm_h = s.positions(out_h, pos, [0, 0, 0, 0], 1337)
dpos = [torch.from_numpy(p).cuda() for p in pos]
out_d = torch.empty(cnt, dtype=torch.float32, device="cuda")
mm_d = s.positions(out_d, dpos, [0, 0, 0, 0], 1337)
assert mm_h == mm_d
assert np.array_equal(out_h.view(np.uint32), out_d.cpu().numpy().view(np.uint32))
# unaligned device views take the scalar load/store path
out_u = torch.empty(cnt + 1, dtype=torch.float32, device="cuda")[1:]
upos = [torch.cat([p.new_zeros(1), p])[1:] for p in dpos]
s.positions(out_u, upos, [0, 0, 0, 0], 1337)
assert torch.equal(out_u, out_d)
print("ok")
''' % helpers.ROOT
    result = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert result.returncode == 0 and result.stdout.strip().endswith("ok"), result.stderr[-2000:]


def test_slabs_reproduce_the_whole_grid_bit_for_bit(gpu):
    """Size-independent property behind multi-GPU sharding: any split of the outermost axis into slabs, generated with
    global integer indices, concatenates to exactly the single-call grid -- at BASELINE size for config 1 (256^3)."""
    import torch
    lib = gpu.lib
    g = gpu.Generator(gpu.TREES["c1_simplex"])
    n = 256
    whole = torch.empty(n ** 3, dtype=torch.float32, device="cuda")
    mm_whole = g.grid3d(whole, 0, 0, 0, n, n, n, 1.0, 1.0, 1.0, 1337)
    parts = torch.empty(n ** 3, dtype=torch.float32, device="cuda")
    acc = torch.zeros(2, dtype=torch.int32, device="cuda")
    cuts = [0, 1, 37, 128, 129, 255, 256]
    for z0, z1 in zip(cuts[:-1], cuts[1:]):
        rc = lib.fnb200GenUniformGrid3D(g.handle, parts[z0 * n * n:].data_ptr(), 0.0, 0.0, 0.0, n, n, n, 1.0, 1.0, 1.0, 1337, z0, z1, acc.data_ptr(), None)
        assert rc == 0, gpu.last_error()
    mm = np.zeros(2, np.float32)
    assert lib.fnb200MinMaxRead(acc.data_ptr(), None, mm.ctypes.data) == 0
    assert torch.equal(whole, parts)
    assert (float(mm[0]), float(mm[1])) == mm_whole
    # a float-offset slab (what the unmodified reference API offers) is NOT guaranteed identical; the integer-origin entry is


def test_full_size_terrain_slabs_and_range(gpu):
    """BASELINE config 4 at full size on one GPU: 8 z-slabs of 128 == one 1024^3 call; output stays in Remap/Terrace's [0,1]."""
    import torch
    lib = gpu.lib
    g = gpu.Generator(gpu.TREES["c4_terrain"])
    n = 1024
    whole = torch.empty(n ** 3, dtype=torch.float32, device="cuda")
    mm_whole = g.grid3d(whole, 0, 0, 0, n, n, n, 1.0, 1.0, 1.0, 1337)
    assert 0.0 <= mm_whole[0] < mm_whole[1] <= 1.0
    slab = torch.empty(n * n * 128, dtype=torch.float32, device="cuda")
    acc = torch.zeros(2, dtype=torch.int32, device="cuda")
    for r in range(8):
        assert lib.fnb200GenUniformGrid3D(g.handle, slab.data_ptr(), 0.0, 0.0, 0.0, n, n, n, 1.0, 1.0, 1.0, 1337, r * 128, (r + 1) * 128, acc.data_ptr(), None) == 0
        assert torch.equal(slab, whole[r * 128 * n * n:(r + 1) * 128 * n * n]), r
    mm = np.zeros(2, np.float32)
    lib.fnb200MinMaxRead(acc.data_ptr(), None, mm.ctypes.data)
    assert (float(mm[0]), float(mm[1])) == mm_whole
    assert float(whole.min()) == mm_whole[0] and float(whole.max()) == mm_whole[1]


def test_relaxed_math_mode_within_tolerance(gpu, ref, flavour):
    """FMA contraction + approximate rsqrt/rcp: >= 95 % of points within 2e-5 of the strict oracle and every point within
    8e-3 (a contracted skew can flip a simplex lattice decision on ~0.2 % of points, |d| up to 5.3e-3 measured -- the same band the
    reference's own relaxed CPU builds show against its strict build, tests/test_oracle.py). C3b is excluded: White hashes
    coordinate BIT PATTERNS, so it is only meaningful in strict mode."""
    gpu.set_math_mode(flavour.mode | 1)
    try:
        for name in ("c1_simplex", "c3a_cellular_distance", "c3c_cellular_lookup_gradient", "c4_terrain"):
            tree = gpu.TREES[name]
            case = dict(kind="grid3d", off=(0.0, 0.0, 0.0), n=(96, 96, 48), step=(1.7, 1.7, 1.7), seed=1337, tree=tree)
            got, _ = cases.run_case(helpers.ProductNode(gpu, tree), case)
            want, _ = cases.run_case(helpers.RefNode(ref, tree), case)
            d = np.abs(got.astype(np.float64) - want)
            assert (d <= TOL).mean() >= 0.95, (name, float((d <= TOL).mean()))
            assert d.max() <= 8e-3 or name == "c3c_cellular_lookup_gradient", (name, float(d.max()))
    finally:
        gpu.set_math_mode(flavour.mode)


def test_philox_fill_is_reproducible_and_uniform(gpu):
    import torch
    lib = gpu.lib
    a = torch.empty(1 << 20, dtype=torch.float32, device="cuda")
    b = torch.empty(1 << 19, dtype=torch.float32, device="cuda")
    assert lib.fnb200FillUniformPhilox(a.data_ptr(), a.numel(), 0, 1337, 2, -1000.0, 1000.0, None) == 0
    assert lib.fnb200FillUniformPhilox(b.data_ptr(), b.numel(), 1 << 19, 1337, 2, -1000.0, 1000.0, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(a[1 << 19:], b)  # counter-based: any rank regenerates any index range
    assert -1000.0 <= float(a.min()) and float(a.max()) < 1000.0 and abs(float(a.mean())) < 3.0


def test_bounded_inv_sqrt_equals_ieee_on_its_whole_domain(gpu):
    """The Cellular jitter vectors use a check-free Newton sequence for 1/sqrt (fn2::inv_sqrt_bounded). Its arguments are sums
    of squares of half-odd integers, all inside [0.5, 2^22); every one of the 193 M floats of [0.5, 2^23) must give the bits
    of the IEEE `1 / sqrt` sequence the strict oracle computes."""
    assert gpu.lib.fnb200SelfTestInvSqrt(0.5, float(2 ** 23)) == 0


def test_lattice_aligned_grids(gpu, ref):
    """Integer coordinates at feature scale 1 put every point on a lattice vertex: offsets and gradient dots are exact zeros
    (whose sign is one of the places the two gradient flavours differ: a zero table weight contributes a signed zero) and
    every floor / round-to-nearest decision sits on a tie. Bits must still match the flavour's own reference build."""
    def build(lib, node_id):
        t = trees.Tree(lib)
        h = t.new(node_id)
        t.set_float(h, "Feature Scale", 1.0)
        return t, h
    for node_id in (6, 7, 8, 9, 10, 11):            # Simplex, SuperSimplex, Perlin, Value, CellularValue, CellularDistance
        for dim, n in ((2, 24), (3, 12), (4, 6)):
            outs = []
            for lib in (gpu.lib, ref):
                t, h = build(lib, node_id)
                out = np.empty(n ** dim, np.float32)
                if dim == 2:
                    lib.fnGenUniformGrid2D(h, out.ctypes.data, -7.0, -5.0, n, n, 1.0, 1.0, 1337, None)
                elif dim == 3:
                    lib.fnGenUniformGrid3D(h, out.ctypes.data, -7.0, -5.0, -3.0, n, n, n, 1.0, 1.0, 1.0, 1337, None)
                else:
                    lib.fnGenUniformGrid4D(h, out.ctypes.data, -3.0, -2.0, -1.0, 0.0, n, n, n, n, 1.0, 1.0, 1.0, 1.0, 1337, None)
                outs.append(out)
                t.close()
            assert np.array_equal(bits(outs[0]), bits(outs[1])), (node_id, dim, helpers.compare(outs[0], outs[1]))


def test_random_trees_vs_strict_oracle(gpu, ref):
    """Fuzz the lowering: random trees over all node types (random parameters, hybrids as generators, dimension changes nested
    in warps and fractals, shared sub-trees) must reproduce the oracle bit for bit through the C ABI in 2D, 3D and 4D."""
    from tests.fuzz import TreeFuzzer, same
    fuzzer = TreeFuzzer(4242)
    failures = []
    for index in range(100):
        text = fuzzer.tree(depth=3).encode()
        a, b = helpers.ProductNode(gpu, text), helpers.RefNode(ref, text)
        for dim, got, want in ((2, a.grid2d(-3.3, 2.1, 48, 40, 1.9, 1.7, index)[0], b.grid2d(-3.3, 2.1, 48, 40, 1.9, 1.7, index)[0]),
                               (3, a.grid3d(-3.3, 2.1, 0.7, 16, 14, 12, 1.9, 1.7, 2.3, index)[0], b.grid3d(-3.3, 2.1, 0.7, 16, 14, 12, 1.9, 1.7, 2.3, index)[0]),
                               (4, a.grid4d(-3.3, 2.1, 0.7, -1.9, 8, 7, 6, 5, 1.9, 1.7, 2.3, 1.1, index)[0], b.grid4d(-3.3, 2.1, 0.7, -1.9, 8, 7, 6, 5, 1.9, 1.7, 2.3, 1.1, index)[0])):
            if not same(got, want):
                failures.append((index, dim, text, helpers.compare(got, want)))
        if index % 4 == 0:                          # the other two entry kinds: scattered positions and the torus mapping
            rng = np.random.default_rng(index)
            pos = [rng.uniform(-300, 300, 1537).astype(np.float32) for _ in range(3)]
            if not same(a.positions(pos, [0.5, -1.5, 2.5], index)[0], b.positions(pos, [0.5, -1.5, 2.5], index)[0]):
                failures.append((index, "positions3d", text))
            if not same(a.tileable2d(40, 24, 1.3, 0.9, index)[0], b.tileable2d(40, 24, 1.3, 0.9, index)[0]):
                failures.append((index, "tileable2d", text))
    assert not failures, failures[:5]


def test_concurrent_calls_from_threads(gpu):
    """The reference's Gen* methods are const and thread-safe; here every calling thread owns its streams, staging buffers
    and min/max accumulator. Eight threads hammer two shared generators (host and device destinations, first use compiles
    concurrently) and every result must equal the single-threaded one."""
    import threading
    import torch
    gens = [gpu.Generator(gpu.TREES["c4_terrain"]), gpu.Generator(gpu.TREES["c3a_cellular_distance"])]
    shapes = [(40, 24, 16), (33, 17, 9)]
    expected = {}
    for gi, g in enumerate(gens):
        for si, n in enumerate(shapes):
            out = np.empty(n[0] * n[1] * n[2], np.float32)
            expected[gi, si] = (out, g.grid3d(out, 1.5, -2.5, 3.5, *n, 0.7, 0.7, 0.7, 99))
    errors = []

    def worker(tid):
        try:
            for it in range(40):
                gi, si = (tid + it) % 2, (tid // 2 + it) % 2
                n = shapes[si]
                want, want_mm = expected[gi, si]
                if (tid + it) % 3 == 0:
                    dev = torch.empty(want.size, dtype=torch.float32, device="cuda")
                    mm = gens[gi].grid3d(dev, 1.5, -2.5, 3.5, *n, 0.7, 0.7, 0.7, 99)
                    got = dev.cpu().numpy()
                else:
                    got = np.empty(want.size, np.float32)
                    mm = gens[gi].grid3d(got, 1.5, -2.5, 3.5, *n, 0.7, 0.7, 0.7, 99)
                if mm != want_mm or not np.array_equal(bits(got), bits(want)):
                    errors.append((tid, it, gi, si))
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


def test_packed_leaves_equal_scalar_leaves(gpu):
    """Every leaf of the device library evaluated two points per value (packed FP32: FFMA2 / FMUL2 / FADD2 for the arithmetic,
    lane by lane for everything else) must give the bits of the one-point-per-value instantiation."""
    n = 1 << 16
    failures = []
    for which in range(24):
        for step in (0.37, 0.011):
            a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
            assert gpu.lib.fnb200SelfTestLeaf(0, which, n, 1337, step, a.ctypes.data) == 0
            assert gpu.lib.fnb200SelfTestLeaf(1, which, n, 1337, step, b.ctypes.data) == 0
            if not np.array_equal(bits(a), bits(b)):
                failures.append((which, step, helpers.compare(b, a)))
    assert not failures, failures


@pytest.fixture
def force_packed():
    """Small test grids normally run one point per thread; this makes them take the code path of large calls (4 points per
    thread) with two points per value (packed FP32) in every dimension (opt-in in production: FASTNOISE2_B200_PACK)."""
    saved = {k: os.environ.get(k) for k in ("FASTNOISE2_B200_PACK", "FASTNOISE2_B200_FILL")}
    os.environ["FASTNOISE2_B200_PACK"] = "2"
    os.environ["FASTNOISE2_B200_FILL"] = "0"
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def test_packed_path_every_node_and_random_trees(gpu, ref, flavour, force_packed):
    """The packed-FP32 instantiation of the whole node library (opt-in, FASTNOISE2_B200_PACK): every node type in 2D/3D/4D with
    constant and generated hybrids, plus 16 random trees, bit for bit against the oracle. Table-gradient flavour only: every kernel
    of this test is a fresh NVRTC compile of the slow-to-compile packed form (the single largest cost of the GPU suite), the generic
    flavour of the packed form is covered on the host (test_packed_generated_code_on_host_matches_oracle) and by the leaf self-test."""
    from tests.fuzz import TreeFuzzer, same
    if flavour.name != "tables":
        pytest.skip("packed form: table flavour on the GPU, both flavours on the host emulation")
    failures = []
    for node_id in range(47):
        for dim, n in ((2, 64), (3, 20), (4, 9)):
            a, b = trees.Tree(gpu.lib), trees.Tree(ref)
            ha, hb = a.wire(node_id, node_id % 2 == 1), b.wire(node_id, node_id % 2 == 1)
            outs = []
            for lib, h in ((gpu.lib, ha), (ref, hb)):
                out = np.empty(n ** dim, np.float32)
                if dim == 2:
                    lib.fnGenUniformGrid2D(h, out.ctypes.data, -3.3, 2.1, n, n, 1.9, 1.9, 1337, None)
                elif dim == 3:
                    lib.fnGenUniformGrid3D(h, out.ctypes.data, -3.3, 2.1, 0.7, n, n, n, 1.9, 1.9, 1.9, 1337, None)
                else:
                    lib.fnGenUniformGrid4D(h, out.ctypes.data, -3.3, 2.1, 0.7, -1.9, n, n, n, n, 1.9, 1.9, 1.9, 1.9, 1337, None)
                outs.append(out)
            if not same(outs[0], outs[1]):
                failures.append((node_id, dim, helpers.compare(outs[0], outs[1])))
            a.close()
            b.close()
    fuzzer = TreeFuzzer(777)
    for index in range(16):
        text = fuzzer.tree(depth=3).encode()
        a, b = helpers.ProductNode(gpu, text), helpers.RefNode(ref, text)
        for dim, got, want in ((2, a.grid2d(-3.3, 2.1, 48, 40, 1.9, 1.7, index)[0], b.grid2d(-3.3, 2.1, 48, 40, 1.9, 1.7, index)[0]),
                               (3, a.grid3d(-3.3, 2.1, 0.7, 16, 14, 12, 1.9, 1.7, 2.3, index)[0], b.grid3d(-3.3, 2.1, 0.7, 16, 14, 12, 1.9, 1.7, 2.3, index)[0]),
                               (4, a.grid4d(-3.3, 2.1, 0.7, -1.9, 8, 7, 6, 5, 1.9, 1.7, 2.3, 1.1, index)[0], b.grid4d(-3.3, 2.1, 0.7, -1.9, 8, 7, 6, 5, 1.9, 1.7, 2.3, 1.1, index)[0])):
            if not same(got, want):
                failures.append((index, dim, text, helpers.compare(got, want)))
        if index % 4 == 0:
            rng = np.random.default_rng(index)
            pos = [rng.uniform(-300, 300, 1537).astype(np.float32) for _ in range(3)]
            if not same(a.positions(pos, [0.5, -1.5, 2.5], index)[0], b.positions(pos, [0.5, -1.5, 2.5], index)[0]):
                failures.append((index, "positions3d", text))
            if not same(a.tileable2d(40, 24, 1.3, 0.9, index)[0], b.tileable2d(40, 24, 1.3, 0.9, index)[0]):
                failures.append((index, "tileable2d", text))
    assert not failures, failures[:5]


@pytest.fixture
def force_large_call_path():
    """Small test grids normally run one point per thread; FASTNOISE2_B200_FILL=0 makes them take the code path of large calls,
    FASTNOISE2_B200_RAGGED=2 keeps the any-width four-point form for heavy trees too (production picks it by code size)."""
    saved = {k: os.environ.get(k) for k in ("FASTNOISE2_B200_FILL", "FASTNOISE2_B200_RAGGED")}
    os.environ["FASTNOISE2_B200_FILL"] = "0"
    os.environ["FASTNOISE2_B200_RAGGED"] = "2"
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def test_any_width_rows_keep_four_points_per_thread(gpu, ref, force_large_call_path):
    """Rows whose width 4 does not divide (chunk sizes such as the (128+2)^3 of tools/NodeEditor/MeshNoisePreview.cpp:446 as
    uniform grids) run four consecutive output points per thread, groups straddling row / plane / volume
    ends, partial last group: bit-identical to the oracle in 2D, 3D and 4D, whole calls and row slabs."""
    launches = gpu.lib.fnb200LaunchCount()
    for name in ("c4_terrain", "c2_fbm5_perlin"):
        a, b = helpers.ProductNode(gpu, gpu.TREES[name]), helpers.RefNode(ref, gpu.TREES[name])
        for n in [(5, 3, 2), (7, 9, 3), (130, 5, 3), (33, 2, 2), (6, 1, 1), (9, 1, 5)]:
            got, gmm = a.grid3d(0.5, -0.25, 3.0, *n, 0.11, 0.13, 0.17, 1337)
            want, wmm = b.grid3d(0.5, -0.25, 3.0, *n, 0.11, 0.13, 0.17, 1337)
            assert np.array_equal(bits(got), bits(want)), (name, n)
            assert tuple(gmm) == tuple(wmm)
        for n in [(5, 1), (7, 9), (131, 17), (6, 5)]:
            got, gmm = a.grid2d(-3.0, 2.0, *n, 0.09, 0.07, 5)
            want, wmm = b.grid2d(-3.0, 2.0, *n, 0.09, 0.07, 5)
            assert np.array_equal(bits(got), bits(want)), (name, n)
            assert tuple(gmm) == tuple(wmm)
        for n in [(5, 3, 2, 2), (7, 2, 3, 2), (9, 1, 1, 3)]:
            got, gmm = a.grid4d(0.0, 1.0, 2.0, 3.0, *n, 0.19, 0.17, 0.13, 0.11, 5)
            want, wmm = b.grid4d(0.0, 1.0, 2.0, 3.0, *n, 0.19, 0.17, 0.13, 0.11, 5)
            assert np.array_equal(bits(got), bits(want)), (name, n)
            assert tuple(gmm) == tuple(wmm)
    # several staging chunks (each starts at its own row of the grid) and a device-resident destination
    os.environ["FASTNOISE2_B200_STAGING_MB"] = "1"
    try:
        a, b = helpers.ProductNode(gpu, gpu.TREES["c1_simplex"]), helpers.RefNode(ref, gpu.TREES["c1_simplex"])
        got, gmm = a.grid3d(0.5, -0.25, 3.0, 130, 70, 61, 0.11, 0.13, 0.17, 7)
        want, wmm = b.grid3d(0.5, -0.25, 3.0, 130, 70, 61, 0.11, 0.13, 0.17, 7)
        assert np.array_equal(bits(got), bits(want)) and tuple(gmm) == tuple(wmm)
    finally:
        os.environ.pop("FASTNOISE2_B200_STAGING_MB", None)
    import torch
    dev = torch.empty(130 * 70 * 61, dtype=torch.float32, device="cuda")
    mm = gpu.Generator(gpu.TREES["c1_simplex"]).grid3d(dev, 0.5, -0.25, 3.0, 130, 70, 61, 0.11, 0.13, 0.17, 7)
    assert np.array_equal(bits(dev.cpu().numpy()), bits(want)) and tuple(mm) == tuple(wmm)
    assert gpu.lib.fnb200LaunchCount() > launches
    # the any-width entry is what ran: its text is what fnb200GetKernelSource returns under FASTNOISE2_B200_PPT=5
    os.environ["FASTNOISE2_B200_PPT"] = "5"
    try:
        assert "rows of any width" in gpu.Generator(gpu.TREES["c4_terrain"]).kernel_source(3)
    finally:
        os.environ.pop("FASTNOISE2_B200_PPT", None)


def test_chunk_sized_odd_width_call_at_full_size(gpu, ref):
    """130^3 (a mesh-preview chunk with its one-voxel apron) is large enough to take the four-point path on its own."""
    tree = gpu.TREES["c4_terrain"]
    a, b = helpers.ProductNode(gpu, tree), helpers.RefNode(ref, tree)
    got, gmm = a.grid3d(-65.0, -65.0, -65.0, 130, 130, 130, 1.0, 1.0, 1.0, 1337)
    want, wmm = b.grid3d(-65.0, -65.0, -65.0, 130, 130, 130, 1.0, 1.0, 1.0, 1337)
    assert np.array_equal(bits(got), bits(want))
    assert tuple(gmm) == tuple(wmm)
