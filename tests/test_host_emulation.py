"""The lowering, checked on the CPU. The CUDA source that lower.cpp generates for a tree (plus the device node library it
includes) is compiled with the HOST compiler over tests/host_emulation/cuda_host_shim.h and run "thread" by "thread"; the
values must equal the strict oracle's bit for bit. This exercises the emitter and the scalar and packed instantiations of
the node library for random trees without a GPU. It is a test harness around product-generated code, not a product path:
nothing under fastnoise2_b200/ can evaluate noise on the CPU.

    python tests/test_host_emulation.py 2000 [seed base] [tree depth]   # offline fuzz with that many random trees per flavour
    FN2_FUZZ_GRIDS=far ...      grids far from the origin (big cell coordinates; Convert<int> overflow shows up in 4D)
    FN2_FUZZ_ENTRIES=1 ...      every tree also through the tileable and position-array entry kernels
    FN2_FUZZ_WILD=0.25 ...      that share of the parameters leaves its sensible / valid range (NaN, inf, out-of-range enums)
    FN2_FUZZ_SETTERS=1 ...      random reflection-setter sequences (in- and out-of-range values) instead of encoded trees
    FASTNOISE2_B200_PACK=2 FASTNOISE2_B200_FILL=0 ...   the packed-FP32 form
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import helpers  # noqa: E402
from tests.fuzz import TreeFuzzer, same  # noqa: E402

# x sizes are multiples of 4: fnb200GetKernelSource hands out the 4-points-per-thread variant, whose threads own x-adjacent
# points of ONE row (the library picks a smaller variant for other widths at launch time)
GRIDS = {2: dict(off=(-3.3, 2.1), n=(20, 13), step=(1.9, 1.7)),
         3: dict(off=(-3.3, 2.1, 0.7), n=(8, 7, 5), step=(1.9, 1.7, 2.3)),
         4: dict(off=(-3.3, 2.1, 0.7, -1.9), n=(4, 5, 3, 3), step=(1.9, 1.7, 2.3, 1.1))}


_BUILDS = [0]
KIND_TILEABLE, KIND_POS = 5, 10


def build_on_host(fn, text, kind, workdir, tag):
    """Lower `text` for one entry kind, compile the generated CUDA with g++ over the stand-in; returns (library, points per thread)."""
    _BUILDS[0] += 1
    tag = "%s_%d" % (tag, _BUILDS[0])  # dlopen() hands back an already loaded library for a path it has seen
    gen = fn.Generator(text)
    source = gen.kernel_source(kind)
    params = gen.kernel_params(kind)
    ppt = int(source.split(" point(s) per thread")[0].rsplit(" ", 1)[1])
    entry = 1 if kind == KIND_TILEABLE else (2 if kind > KIND_POS else 0)
    path = os.path.join(workdir, "tree_%s.cpp" % tag)
    with open(path, "w") as f:
        f.write('#include "cuda_host_shim.h"\n' + source + '\n#define FN2_HOST_ENTRY %d\n#include "driver.inl"\n' % entry)
    lib_path = os.path.join(workdir, "tree_%s.so" % tag)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fno-fast-math", "-fwrapv", "-w", "-fPIC", "-shared",
                           "-I" + os.path.join(ROOT, "tests", "host_emulation"), "-I" + os.path.join(ROOT, "fastnoise2_b200", "csrc", "device"), path, "-o", lib_path])
    return ctypes.CDLL(lib_path), (ppt, params)


def run_on_host(fn, text, dims, seed, workdir, tag):
    """dims-D uniform grid of GRIDS[dims] through the generated entry kernel, on the host."""
    lib, (ppt, params) = build_on_host(fn, text, dims, workdir, tag)
    g = GRIDS[dims]
    out = np.zeros(int(np.prod(g["n"])), np.float32)
    lib.fn2_host_grid(out.ctypes.data_as(ctypes.c_void_p), dims, (ctypes.c_float * dims)(*g["off"]), (ctypes.c_int * dims)(*g["n"]),
                      (ctypes.c_float * dims)(*g["step"]), seed, ppt, params.ctypes.data_as(ctypes.c_void_p), len(params))
    return out


def run_tileable_on_host(fn, text, nx, ny, sx, sy, seed, workdir, tag):
    lib, (ppt, params) = build_on_host(fn, text, KIND_TILEABLE, workdir, tag)
    out = np.zeros(nx * ny, np.float32)
    lib.fn2_host_tileable(out.ctypes.data_as(ctypes.c_void_p), nx, ny, ctypes.c_float(sx), ctypes.c_float(sy), seed, ppt, params.ctypes.data_as(ctypes.c_void_p), len(params))
    return out


def run_positions_on_host(fn, text, pos, offsets, seed, workdir, tag):
    dims = len(pos)
    lib, (ppt, params) = build_on_host(fn, text, KIND_POS + dims, workdir, tag)
    out = np.zeros(len(pos[0]), np.float32)
    ptrs = [p.ctypes.data_as(ctypes.c_void_p) for p in pos] + [None] * (4 - dims)
    lib.fn2_host_positions(out.ctypes.data_as(ctypes.c_void_p), dims, ctypes.c_longlong(len(pos[0])), *ptrs, (ctypes.c_float * dims)(*offsets), seed, ppt, params.ctypes.data_as(ctypes.c_void_p), len(params))
    return out


def oracle_values(node, dims, seed):
    g = GRIDS[dims]
    call = {2: node.grid2d, 3: node.grid3d, 4: node.grid4d}[dims]
    return call(*g["off"], *g["n"], *g["step"], seed)[0]


def test_non_finite_values_follow_the_x86_oracle_on_host(fn, ref, flavour, tmp_path):
    """NaN / inf / signed zeros produced INSIDE a tree take the oracle's path through Min / Max / MaxSmooth / the fractals."""
    from tests.cases import NON_FINITE_TREES
    failures = []
    for name, text in NON_FINITE_TREES.items():
        for dims in (2, 3):
            got = run_on_host(fn, text, dims, 146, str(tmp_path), "nonfinite")
            want = oracle_values(helpers.RefNode(ref, text), dims, 146)
            if not same(got, want):
                failures.append((name, dims, helpers.compare(got, want)))
    assert not failures, failures


def test_convert_overflow_regime_matches_the_saturating_port_on_host(fn, port, flavour, tmp_path, monkeypatch):
    """Cell coordinates beyond +-2^31: the product's Convert<int> saturates; the C restatement in its saturating mode is the oracle."""
    from tests.cases import CONVERT_OVERFLOW_TREES, FAR_GRID_4D
    monkeypatch.setitem(GRIDS, 4, FAR_GRID_4D)
    port.fn2oSetConvertMode(1)
    try:
        for seed, text in CONVERT_OVERFLOW_TREES:
            got = run_on_host(fn, text, 4, seed, str(tmp_path), "overflow")
            assert same(got, oracle_values(helpers.PortNode(port, text), 4, seed)), text
    finally:
        port.fn2oSetConvertMode(0)


def test_out_of_range_enum_values_follow_the_reference_switches_on_host(fn, ref, tmp_path):
    """Enum members outside their range take the `default:` label / fall-through / no-case path of the reference's switch."""
    from tests.cases import out_of_range_enum_trees
    failures = []
    for name, text in out_of_range_enum_trees().items():
        for dims in (2, 3, 4) if "rotate" in name or "remove" in name else (3,):
            got = run_on_host(fn, text, dims, 31, str(tmp_path), "enum")
            want = oracle_values(helpers.RefNode(ref, text), dims, 31)
            if not same(got, want):
                failures.append((name, dims, helpers.compare(got, want)))
    assert not failures, failures


class _Handles:
    """The same node built through the reflection C ABI in the product and in an oracle library."""

    def __init__(self, fn, ref_lib, node_id):
        self.fn, self.ref_lib = fn, ref_lib
        self.ours = [fn.Generator.from_metadata(node_id)]
        self.theirs = [ref_lib.fnNewFromMetadata(node_id, 0xFFFFFFFF)]

    def leaf(self, node_id, scale):
        self.ours.append(self.fn.Generator.from_metadata(node_id))
        self.theirs.append(self.ref_lib.fnNewFromMetadata(node_id, 0xFFFFFFFF))
        for lib, handle in ((self.fn.lib, self.ours[-1].handle), (self.ref_lib, self.theirs[-1])):
            assert lib.fnSetVariableFloat(handle, 0, ctypes.c_float(scale))
        return self.ours[-1].handle, self.theirs[-1]

    def both(self, name, *args):
        """Calls fn<name> on the root of both trees; args may be (ours, theirs) pairs."""
        a = getattr(self.fn.lib, name)(self.ours[0].handle, *[x[0] if isinstance(x, tuple) else x for x in args])
        b = getattr(self.ref_lib, name)(self.theirs[0], *[x[1] if isinstance(x, tuple) else x for x in args])
        assert bool(a) == bool(b), (name, args, a, b)

    def close(self):
        for h in self.theirs:
            self.ref_lib.fnDeleteNodeRef(h)


def setter_fuzz(fn, ref_lib, count, seed, workdir):
    """Random sequences of fnSetVariableFloat / fnSetVariableIntEnum (values in and out of range) / fnSetHybridFloat /
    fnSetHybridNodeLookup on a random node type, applied to the product and to the oracle; both must accept or reject alike and
    the product's generated kernel (run on the host) must reproduce the oracle's values: setter side effects, enum defaults."""
    import json
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "metadata.json")))
    rng = np.random.default_rng(seed)
    failures = []
    for index in range(count):
        m = meta[int(rng.integers(len(meta)))]
        tree = _Handles(fn, ref_lib, m["id"])
        for i in range(len(m["lookups"])):
            if m["id"] in (19, 20):                   # the warp fractals need a DomainWarp source
                warp = _Handles(fn, ref_lib, 18)
                warp.both("fnSetNodeLookup", 0, warp.leaf(6, 41.0))
                tree.ours += warp.ours
                tree.theirs += warp.theirs
                tree.both("fnSetNodeLookup", i, (warp.ours[0].handle, warp.theirs[0]))
            else:
                tree.both("fnSetNodeLookup", i, tree.leaf(6 + i % 3, 37.0 + 11 * i))
        log = []
        for _ in range(int(rng.integers(2, 9))):
            what = int(rng.integers(3))
            if what == 0 and m["variables"]:
                i = int(rng.integers(len(m["variables"])))
                v = m["variables"][i]
                if v["type"] == 0:
                    x = float(np.float32(rng.choice([rng.uniform(-3, 3), rng.uniform(0.1, 80), 0.0, -1.0, 1.0])))
                    if "Scale" in v["name"] and x == 0.0:
                        x = 2.5                       # a zero scale is an infinite frequency: NaN positions
                    tree.both("fnSetVariableFloat", i, ctypes.c_float(x))
                else:
                    x = int(rng.integers(1, 5)) if "ctave" in v["name"] else int(rng.integers(-1, 7))
                    tree.both("fnSetVariableIntEnum", i, x)
                log.append((v["name"], v["dim"], x))
            elif what == 1 and m["hybrids"]:
                i = int(rng.integers(len(m["hybrids"])))
                x = float(np.float32(rng.choice([rng.uniform(-2, 2), rng.uniform(0.1, 3), 0.0, 0.5])))
                tree.both("fnSetHybridFloat", i, ctypes.c_float(x))
                log.append((m["hybrids"][i]["name"], x))
            elif what == 2 and m["hybrids"]:
                i = int(rng.integers(len(m["hybrids"])))
                leaf, scale = int(rng.choice([6, 8, 9])), float(np.float32(rng.uniform(5, 60)))
                tree.both("fnSetHybridNodeLookup", i, tree.leaf(leaf, scale))
                log.append((m["hybrids"][i]["name"], "node", leaf, scale))
        holder = type("OneTree", (), {"Generator": staticmethod(lambda text, g=tree.ours[0]: g)})
        for dims in (2, 3):
            g = GRIDS[dims]
            lib, (ppt, params) = build_on_host(holder, "", dims, workdir, "setters")
            got = np.zeros(int(np.prod(g["n"])), np.float32)
            lib.fn2_host_grid(got.ctypes.data_as(ctypes.c_void_p), dims, (ctypes.c_float * dims)(*g["off"]), (ctypes.c_int * dims)(*g["n"]),
                              (ctypes.c_float * dims)(*g["step"]), index, ppt, params.ctypes.data_as(ctypes.c_void_p), len(params))
            want, mm = np.empty_like(got), np.zeros(2, np.float32)
            call = ref_lib.fnGenUniformGrid2D if dims == 2 else ref_lib.fnGenUniformGrid3D
            call(tree.theirs[0], want.ctypes.data, *g["off"], *g["n"], *g["step"], index, mm.ctypes.data)
            if not same(got, want):
                failures.append((m["name"], dims, log, helpers.compare(got, want)))
        tree.close()
    return failures


def test_random_setter_sequences_on_host_match_the_oracle(fn, ref, tmp_path):
    """20 random setter sequences (offline: 900 with other seeds, which found the two out-of-range enum cases now in
    tests/cases.py out_of_range_enum_trees)."""
    failures = setter_fuzz(fn, ref, 20, 424242, str(tmp_path))
    assert not failures, failures


def fuzz(fn, ref_lib, count, seed, workdir, label, depth=3):
    fuzzer = TreeFuzzer(seed, wild=float(os.environ.get("FN2_FUZZ_WILD", "0")))
    failures = []
    for index in range(count):
        text = fuzzer.tree(depth=depth).encode()
        for dims in (2, 3, 4):
            got = run_on_host(fn, text, dims, index, workdir, "%s_%d" % (label, dims))
            want = oracle_values(helpers.RefNode(ref_lib, text), dims, index)
            if not same(got, want):
                failures.append((label, index, dims, text, helpers.compare(got, want)))
        if os.environ.get("FN2_FUZZ_ENTRIES"):          # offline runs: also the torus mapping and a ragged position array
            node = helpers.RefNode(ref_lib, text)
            got = run_tileable_on_host(fn, text, 12, 7, 1.3, 0.9, index, workdir, label + "_tile")
            if not same(got, node.tileable2d(12, 7, 1.3, 0.9, index)[0]):
                failures.append((label, index, "tileable", text, {}))
            rng = np.random.default_rng(index)
            pos = [rng.uniform(-300, 300, 203).astype(np.float32) for _ in range(3)]
            got = run_positions_on_host(fn, text, pos, [0.5, -1.5, 2.5], index, workdir, label + "_pos")
            if not same(got, node.positions(pos, [0.5, -1.5, 2.5], index)[0]):
                failures.append((label, index, "positions3d", text, {}))
    return failures


def test_generated_code_on_host_matches_oracle(fn, ref, flavour, tmp_path):
    """One point per value (the default): the five config trees and 12 random trees, in 2D, 3D and 4D."""
    failures = []
    for name in ("c1_simplex", "c2_fbm5_perlin", "c3a_cellular_distance", "c3b_cellular_lookup_white", "c4_terrain", "c5_supersimplex"):
        for dims in (2, 3, 4):
            got = run_on_host(fn, fn.TREES[name], dims, 1337, str(tmp_path), "cfg")
            want = oracle_values(helpers.RefNode(ref, fn.TREES[name]), dims, 1337)
            if not same(got, want):
                failures.append((name, dims, helpers.compare(got, want)))
    failures += fuzz(fn, ref, 12, 20260117, str(tmp_path), flavour.name)
    # the other two entry kinds: torus mapping and scattered positions (ragged count: the tail path of the entry kernel)
    rng = np.random.default_rng(7)
    for name in ("c2_fbm5_perlin", "c4_terrain", "c5_supersimplex"):
        node = helpers.RefNode(ref, fn.TREES[name])
        got = run_tileable_on_host(fn, fn.TREES[name], 24, 10, 1.3, 0.9, 1337, str(tmp_path), "tile")
        if not same(got, node.tileable2d(24, 10, 1.3, 0.9, 1337)[0]):
            failures.append((name, "tileable"))
        for dims in (2, 3, 4):
            pos = [rng.uniform(-300, 300, 203).astype(np.float32) for _ in range(dims)]
            offsets = [0.5, -1.5, 2.5, -3.5][:dims]
            got = run_positions_on_host(fn, fn.TREES[name], pos, offsets, 1337, str(tmp_path), "pos")
            if not same(got, node.positions(pos, offsets, 1337)[0]):
                failures.append((name, "positions", dims))
    assert not failures, failures[:5]


def test_packed_generated_code_on_host_matches_oracle(fn, ref, flavour, tmp_path, monkeypatch):
    """Two points per value (FASTNOISE2_B200_PACK=2): the pack overloads and the templated node library, with the packed FP32
    intrinsics emulated lane by lane, for the config trees and 8 random trees."""
    monkeypatch.setenv("FASTNOISE2_B200_PACK", "2")
    failures = []
    for name in ("c1_simplex", "c3a_cellular_distance", "c4_terrain", "c5_supersimplex"):
        for dims in (2, 3, 4):
            got = run_on_host(fn, fn.TREES[name], dims, 1337, str(tmp_path), "cfgp")
            want = oracle_values(helpers.RefNode(ref, fn.TREES[name]), dims, 1337)
            if not same(got, want):
                failures.append((name, dims, helpers.compare(got, want)))
    failures += fuzz(fn, ref, 8, 424242, str(tmp_path), flavour.name + "_packed")
    assert not failures, failures[:5]


if __name__ == "__main__":
    import tempfile
    import fastnoise2_b200 as product
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    base = int(sys.argv[2]) if len(sys.argv) > 2 else 9000  # fuzzer seed: run several processes with different bases in parallel
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    if os.environ.get("FN2_FUZZ_SETTERS"):                      # random setter sequences instead of random encoded trees
        generic = os.environ["FN2_FUZZ_SETTERS"] == "generic"
        product.set_math_mode(2 if generic else 0)
        for item in setter_fuzz(product, helpers.load_ref("strict_avx2" if generic else "strict_avx512"), total, base, tempfile.mkdtemp()):
            print(item)
        print("setter sequences done", flush=True)
        sys.exit(0)
    if os.environ.get("FN2_FUZZ_GRIDS") == "far":               # far from the origin, negative, coarse: big cell coordinates
        GRIDS.update({2: dict(off=(-91234.5, 70001.25), n=(20, 13), step=(37.7, 53.1)),
                      3: dict(off=(-91234.5, 70001.25, -333.0), n=(8, 7, 5), step=(37.7, 53.1, 0.5)),
                      4: dict(off=(-91234.5, 70001.25, -333.0, 1e6), n=(4, 5, 3, 3), step=(37.7, 53.1, 0.5, 1024.0))})
    bad = []
    for variant, mode, label in (("strict_avx512", 0, "tables"), ("strict_avx2", 2, "generic")):
        product.set_math_mode(mode)
        bad += fuzz(product, helpers.load_ref(variant), total, base + mode, tempfile.mkdtemp(), label, depth)
        print(label, "done,", len(bad), "mismatching (tree, dimension) pairs so far", flush=True)
    for item in bad[:20]:
        print(item)


def test_smaller_points_per_thread_variants_on_host(fn, ref, tmp_path, monkeypatch):
    """The 1-point-per-thread variant (what the library launches for row widths that 4 does not divide, and for calls too small
    to fill the GPU) on odd grid sizes, and the literal-parameter form of the kernels (FASTNOISE2_B200_LITERALS=1)."""
    failures = []
    # "5" = four points per thread over rows of ANY width (groups run over row ends; 9 x 7 x 5 = 315 points leaves a partial last group)
    for ppt, literals, n in (("1", "0", (9, 7, 5)), ("1", "1", (10, 7, 3)), ("4", "1", (12, 7, 3)), ("5", "0", (9, 7, 5)), ("5", "0", (5, 3, 2))):
        monkeypatch.setenv("FASTNOISE2_B200_PPT", ppt)
        monkeypatch.setenv("FASTNOISE2_B200_LITERALS", literals)
        for name in ("c4_terrain", "c3a_cellular_distance"):
            lib, (got_ppt, params) = build_on_host(fn, fn.TREES[name], 3, str(tmp_path), "ppt")
            assert got_ppt == (4 if ppt == "5" else int(ppt)) and (len(params) > 1) == (literals == "0")
            out = np.zeros(int(np.prod(n)), np.float32)
            lib.fn2_host_grid(out.ctypes.data_as(ctypes.c_void_p), 3, (ctypes.c_float * 3)(0.5, -2.0, 3.0), (ctypes.c_int * 3)(*n),
                              (ctypes.c_float * 3)(1.3, 1.1, 0.9), 7, got_ppt, params.ctypes.data_as(ctypes.c_void_p), len(params))
            want = helpers.RefNode(ref, fn.TREES[name]).grid3d(0.5, -2.0, 3.0, *n, 1.3, 1.1, 0.9, 7)[0]
            if not same(out, want):
                failures.append((ppt, name, helpers.compare(out, want)))
    assert not failures, failures


def test_any_width_four_point_groups_on_host(fn, ref, tmp_path, monkeypatch):
    """Four points per thread over rows whose width 4 does not divide (FASTNOISE2_B200_PPT=5 selects that entry form for
    fnb200GetKernelSource): groups straddle row, plane and volume ends; the last group of a call is partial. 2D, 3D and 4D."""
    monkeypatch.setenv("FASTNOISE2_B200_PPT", "5")
    text = fn.TREES["c2_fbm5_perlin"]
    node = helpers.RefNode(ref, text)
    failures = []
    for dims, n in ((2, (7, 9)), (2, (5, 3)), (3, (6, 5, 3)), (3, (13, 3, 3)), (4, (5, 3, 2, 3)), (4, (7, 2, 3, 2))):
        lib, (ppt, params) = build_on_host(fn, text, dims, str(tmp_path), "ragged%d" % dims)
        assert ppt == 4
        off, step = (0.5, -2.0, 3.0, 7.5)[:dims], (0.13, 0.11, 0.09, 0.21)[:dims]
        out = np.zeros(int(np.prod(n)), np.float32)
        lib.fn2_host_grid(out.ctypes.data_as(ctypes.c_void_p), dims, (ctypes.c_float * dims)(*off), (ctypes.c_int * dims)(*n),
                          (ctypes.c_float * dims)(*step), 11, ppt, params.ctypes.data_as(ctypes.c_void_p), len(params))
        want = getattr(node, "grid%dd" % dims)(*off, *n, *step, 11)[0]
        if not same(out, want):
            failures.append((dims, n, helpers.compare(out, want)))
    assert not failures, failures


def _weighted_fractal(fn, ref_lib, node_id):
    """FractalFBm (13) / FractalRidged (15) / DomainWarpFractal (19, 20) at their default octave counts with a weighted strength of 0.5."""
    tree = _Handles(fn, ref_lib, node_id)
    if node_id in (19, 20):
        warp = _Handles(fn, ref_lib, 18)
        warp.both("fnSetNodeLookup", 0, warp.leaf(6, 41.0))
        tree.ours += warp.ours
        tree.theirs += warp.theirs
        tree.both("fnSetNodeLookup", 0, (warp.ours[0].handle, warp.theirs[0]))
    else:
        tree.both("fnSetNodeLookup", 0, tree.leaf(6, 37.0))
    tree.both("fnSetHybridFloat", 1, ctypes.c_float(0.5))  # hybrid 0 = Gain, 1 = Weighted Strength
    return tree


def test_code_shape_knobs_are_bit_neutral_on_host(fn, ref, tmp_path, monkeypatch):
    """The A/B knobs of the lowering (profiles/perf_sweep*.sh) only change the SHAPE of the generated code: the divergent group walk
    (FASTNOISE2_B200_CONVERGENT=0, the round-1 form) and one-body octave loops (FASTNOISE2_B200_ONE_BODY=1 / 2: the first octave runs
    inside the loop, sum starts at -0.0f) must give the oracle's bits -- config trees, and fractals / domain-warp fractals with a
    weighted strength (their per-octave amplitude update moves to the end of the loop body)."""
    failures = []
    n, off, step = (12, 7, 5), (0.5, -2.0, 3.0), (1.3, 1.1, 0.9)

    def check(label, holder, text, reference):
        lib, (ppt, params) = build_on_host(holder, text, 3, str(tmp_path), "shape")
        out = np.zeros(int(np.prod(n)), np.float32)
        lib.fn2_host_grid(out.ctypes.data_as(ctypes.c_void_p), 3, (ctypes.c_float * 3)(*off), (ctypes.c_int * 3)(*n),
                          (ctypes.c_float * 3)(*step), 7, ppt, params.ctypes.data_as(ctypes.c_void_p), len(params))
        want = reference()
        if not same(out, want):
            failures.append((label, helpers.compare(out, want)))

    for one_body, convergent in (("1", "1"), ("0", "0"), ("2", "1")):
        monkeypatch.setenv("FASTNOISE2_B200_ONE_BODY", one_body)
        monkeypatch.setenv("FASTNOISE2_B200_CONVERGENT", convergent)
        for name in ("c4_terrain",):
            node = helpers.RefNode(ref, fn.TREES[name])
            check((one_body, convergent, name), fn, fn.TREES[name], lambda: node.grid3d(*off, *n, *step, 7)[0])
        for node_id in (13, 15, 19, 20):
            tree = _weighted_fractal(fn, ref, node_id)
            holder = type("OneTree", (), {"Generator": staticmethod(lambda text, g=tree.ours[0]: g)})

            def reference():
                want, mm = np.empty(int(np.prod(n)), np.float32), np.zeros(2, np.float32)
                ref.fnGenUniformGrid3D(tree.theirs[0], want.ctypes.data, *off, *n, *step, 7, mm.ctypes.data)
                return want
            assert "weightedStrength" in tree.ours[0].kernel_source(3)
            check((one_body, convergent, node_id), holder, "", reference)
            tree.close()
    assert not failures, failures
