"""The oracles against their pins (no GPU): committed golden vectors and the SURVEY.md 8c checksum table."""
import numpy as np
import pytest

from tests import cases, helpers


def test_strict_ref_reproduces_golden_vectors(ref, flavour):
    outputs, summary, _ = helpers.load_golden(flavour.golden)
    for name, case in cases.GOLDEN_CASES.items():
        out, mm = cases.run_case(helpers.RefNode(ref, case["tree"]), case)
        assert np.array_equal(out.view(np.uint32), outputs[name].view(np.uint32)), name
        assert "%08x" % helpers.fnv1a_words(out) == summary[name]["fnv1a"]
        assert float(mm[0]) == summary[name]["min"] and float(mm[1]) == summary[name]["max"]


def test_strict_ref_reproduces_survey_checksums(ref, fn, flavour):
    """An independent build of the same semantics (the survey's scalar prototype) produced these FNV-1a values."""
    if flavour.name != "generic":
        pytest.skip("the survey's prototype followed the generic gradient code path")
    trees = {**fn.TREES, **cases.DEMO_TREES}
    for name, (kind, fnv, lo, hi) in cases.SURVEY_SMOKE.items():
        out, mm = cases.run_case(helpers.RefNode(ref, trees[name]), dict(cases.survey_call(kind), tree=trees[name]))
        assert "%08x" % helpers.fnv1a_words(out) == fnv, name
        assert abs(float(mm[0]) - lo) < 2e-6 and abs(float(mm[1]) - hi) < 2e-6, name


def test_relaxed_ref_stays_inside_the_tolerance_band(ref, fn, flavour):
    """The timing build (reference default flags: -ffast-math, rsqrt/rcp approximations) differs from the strict one only
    in low bits on smooth nodes -- the band real reference binaries occupy among themselves."""
    relaxed = helpers.load_ref("relaxed_avx512" if flavour.name == "tables" else "relaxed_avx2")
    if relaxed is None:
        pytest.skip("relaxed build missing")
    # measured here (AVX2 / AVX-512 relaxed vs strict): smooth trees agree to 4e-6; Simplex flips a lattice decision on 0.2 %
    # of points (|d| up to 1.6e-3); Cellular moves by up to 1.3e-4 with the 12-bit rsqrt; CellularLookup(White) is
    # unrelated bit soup -- which is why parity is defined against the strict build only
    for name, frac, worst in (("c2_fbm5_perlin_grid2d", 1.0, 2e-5), ("c4_terrain_grid3d", 1.0, 2e-5), ("c5_supersimplex_pos4d", 1.0, 2e-5),
                              ("c1_simplex_grid3d", 0.99, 5e-3), ("c3a_cellular_distance_grid3d", 0.5, 5e-4)):
        case = cases.GOLDEN_CASES[name]
        a, _ = cases.run_case(helpers.RefNode(ref, case["tree"]), case)
        b, _ = cases.run_case(helpers.RefNode(relaxed, case["tree"]), case)
        d = np.abs(a.astype(np.float64) - b)
        assert (d <= 2e-5).mean() >= frac and d.max() <= worst, (name, float((d <= 2e-5).mean()), float(d.max()))


def test_port_oracle_reproduces_golden_vectors(port, flavour):
    """oracle/fn2_oracle.c (plain C restatement) against the committed vectors; runs with or without /root/reference."""
    outputs, summary, _ = helpers.load_golden(flavour.golden)
    checked = 0
    for name, case in cases.GOLDEN_CASES.items():
        node = helpers.PortNode(port, case["tree"])
        if not node.supported:
            continue
        out, mm = cases.run_case(node, case)
        assert np.array_equal(out.view(np.uint32), outputs[name].view(np.uint32)), name
        assert float(mm[0]) == summary[name]["min"] and float(mm[1]) == summary[name]["max"], name
        checked += 1
    assert checked >= 8


def test_test_encoder_matches_known_strings(fn):
    from tests.encode import N
    assert N("FractalFBm", {}, [N("Simplex")]).encode() == "DQkGDA=="                      # FastNoise.h:53
    assert N("FractalFBm", {"Octaves": 5}, [N("Perlin")]).encode() == fn.TREES["c2_fbm5_perlin"]
    assert N("CellularLookup", {}, [N("White")]).encode() == fn.TREES["c3b_cellular_lookup_white"]
    c4 = N("DomainWarpFractalProgressive", {}, [N("DomainWarpGradient", {"Feature Scale": 200.0}, [
        N("Terrace", {"Step Count": 12.0}, [N("Remap", {"Clamp Output": 1}, [N("MaxSmooth", {}, [
            N("FractalFBm", {"Octaves": 5}, [N("Simplex", {"Feature Scale": 150.0})])],
            {"RHS": N("FractalRidged", {"Octaves": 4}, [N("Perlin", {"Feature Scale": 90.0, "Seed Offset": 7})]), "Smoothness": 0.2})],
            {"From Max": 2.0})], {"Smoothness": 0.25})], {"Warp Amplitude": 30.0})])
    assert c4.encode() == fn.TREES["c4_terrain"]


def test_port_oracle_vs_strict_ref_every_node(ref, port):
    """The C restatement against the reference's own sources, every node type it covers, 2D/3D/4D, constant and generated hybrids."""
    from tests.encode import default_tree
    checked = 0
    for hybrid_nodes in (False, True):
        for node_id in list(range(47)) + [16, 17]:  # second visit of the simplex warps: GradientOuterProduct scheme
            tree = default_tree(node_id, hybrid_nodes)
            if node_id in (16, 17) and checked >= 47 * 3 * (1 + hybrid_nodes):
                tree.variables["Vectorization Scheme"] = 1
            text = tree.encode()
            a, b = helpers.PortNode(port, text), helpers.RefNode(ref, text)
            assert a.supported, node_id
            for got, want in ((a.grid2d(-3.3, 2.1, 40, 40, 1.9, 1.9, 1337)[0], b.grid2d(-3.3, 2.1, 40, 40, 1.9, 1.9, 1337)[0]),
                              (a.grid3d(-3.3, 2.1, 0.7, 14, 14, 14, 1.9, 1.9, 1.9, 1337)[0], b.grid3d(-3.3, 2.1, 0.7, 14, 14, 14, 1.9, 1.9, 1.9, 1337)[0]),
                              (a.grid4d(-3.3, 2.1, 0.7, -1.9, 7, 7, 7, 7, 1.9, 1.9, 1.9, 1.9, 1337)[0], b.grid4d(-3.3, 2.1, 0.7, -1.9, 7, 7, 7, 7, 1.9, 1.9, 1.9, 1.9, 1337)[0])):
                same = np.array_equal(got.view(np.uint32), want.view(np.uint32)) or (np.isnan(got).all() and np.isnan(want).all())
                assert same, (node_id, hybrid_nodes, helpers.compare(got, want))
                checked += 1
    assert checked == 49 * 2 * 3


def test_port_oracle_slab_rows_equal_whole_grid(port):
    from fastnoise2_b200.api import TREES
    node = helpers.PortNode(port, TREES["c4_terrain"])
    whole, mm = node.grid3d(0.0, 0.0, 0.0, 12, 10, 9, 3.3, 3.3, 3.3, 1337)
    parts = [node.grid3d(0.0, 0.0, 0.0, 12, 10, 9, 3.3, 3.3, 3.3, 1337, rows=(z0 * 10, (z1 - z0) * 10)) for z0, z1 in ((0, 4), (4, 5), (5, 9))]
    assert np.array_equal(np.concatenate([p[0] for p in parts]).view(np.uint32), whole.view(np.uint32))
    assert min(p[1][0] for p in parts) == mm[0] and max(p[1][1] for p in parts) == mm[1]


def test_random_trees_port_vs_strict_ref(ref, port):
    """Fuzz: random trees over all 47 node types (random parameters, hybrids as generators, shared sub-trees) -- the C
    restatement and the reference's own sources must agree on every output word, in 2D, 3D and 4D."""
    from tests.fuzz import TreeFuzzer, same
    fuzzer = TreeFuzzer(20260117)
    failures = []
    for index in range(300):
        text = fuzzer.tree(depth=3).encode()
        a, b = helpers.PortNode(port, text), helpers.RefNode(ref, text)
        assert a.supported, text
        for dim, got, want in ((2, a.grid2d(-3.3, 2.1, 24, 24, 1.9, 1.7, index)[0], b.grid2d(-3.3, 2.1, 24, 24, 1.9, 1.7, index)[0]),
                               (3, a.grid3d(-3.3, 2.1, 0.7, 10, 10, 10, 1.9, 1.7, 2.3, index)[0], b.grid3d(-3.3, 2.1, 0.7, 10, 10, 10, 1.9, 1.7, 2.3, index)[0]),
                               (4, a.grid4d(-3.3, 2.1, 0.7, -1.9, 6, 6, 6, 6, 1.9, 1.7, 2.3, 1.1, index)[0], b.grid4d(-3.3, 2.1, 0.7, -1.9, 6, 6, 6, 6, 1.9, 1.7, 2.3, 1.1, index)[0])):
            if not same(got, want):
                failures.append((index, dim, text, helpers.compare(got, want)))
    assert not failures, failures[:5]


def test_port_oracle_non_finite_trees_vs_strict_ref(ref, port):
    """NaN / inf / signed zeros born inside a tree (tests/cases.py NON_FINITE_TREES): both oracles take the minps / maxps path."""
    from tests.cases import NON_FINITE_TREES
    from tests.fuzz import same
    for name, text in NON_FINITE_TREES.items():
        a, b = helpers.PortNode(port, text), helpers.RefNode(ref, text)
        assert same(a.grid2d(-3.3, 2.1, 24, 24, 1.9, 1.7, 146)[0], b.grid2d(-3.3, 2.1, 24, 24, 1.9, 1.7, 146)[0]), name
        assert same(a.grid3d(-3.3, 2.1, 0.7, 10, 10, 10, 1.9, 1.7, 2.3, 146)[0], b.grid3d(-3.3, 2.1, 0.7, 10, 10, 10, 1.9, 1.7, 2.3, 146)[0]), name


def test_port_oracle_out_of_range_enums_vs_strict_ref(ref, port):
    """Enum members outside their range (tests/cases.py out_of_range_enum_trees): the restatement takes the same switch paths."""
    from tests.cases import out_of_range_enum_trees
    from tests.fuzz import same
    for name, text in out_of_range_enum_trees().items():
        a, b = helpers.PortNode(port, text), helpers.RefNode(ref, text)
        assert a.supported, name
        assert same(a.grid3d(-3.3, 2.1, 0.7, 10, 10, 10, 1.9, 1.7, 2.3, 31)[0], b.grid3d(-3.3, 2.1, 0.7, 10, 10, 10, 1.9, 1.7, 2.3, 31)[0]), name
        assert same(a.grid4d(-3.3, 2.1, 0.7, -1.9, 6, 6, 6, 6, 1.9, 1.7, 2.3, 1.1, 31)[0], b.grid4d(-3.3, 2.1, 0.7, -1.9, 6, 6, 6, 6, 1.9, 1.7, 2.3, 1.1, 31)[0]), name


def test_threaded_tileable_timing_driver_tracks_the_reference_call(ref, fn):
    """bench.py's CPU baseline for the tileable config splits the rows across threads through the reference's GenPositionArray4D
    (GenTileable2D has no slab form). It is a timing driver -- its positions come from libm, not FS::Cos / FS::Sin -- so it must
    follow fnGenTileable2D closely, not bit for bit, and must not depend on the thread count."""
    if not hasattr(ref, "fn2refTimedTileableMT"):
        pytest.skip("oracle/_ref built before the threaded tileable driver existed")
    node = helpers.RefNode(ref, fn.TREES["c2_fbm5_perlin"])
    nx, ny = 96, 40
    want = np.empty(nx * ny, np.float32)
    ref.fnGenTileable2D(node.handle, want.ctypes.data, nx, ny, 1.0, 0.5, 1337, None)
    outs = []
    for threads in (1, 3, 7):
        got = np.empty(nx * ny, np.float32)
        mm = np.zeros(2, np.float32)
        assert ref.fn2refTimedTileableMT(node.handle, got.ctypes.data, nx, ny, 1.0, 0.5, 1337, threads, mm.ctypes.data) > 0
        assert np.abs(got - want).max() < 2e-2 and mm[0] == got.min() and mm[1] == got.max()
        outs.append(got)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
