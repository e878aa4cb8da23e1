"""Shared case tables: the golden cases (small enough to commit) and the SURVEY.md 8c smoke table."""
import numpy as np

from fastnoise2_b200.api import TREES

DEMO_TREES = {
    # shipped with the reference's Node Editor (tools/NodeEditor/util/DemoNodeTrees.inl:5-6): written by the REAL encoder
    "mountain_terrain": "E@BBZEG@BD8JFgIECArXIzwECiQIw/UoPwkuAAE@BJDQAH@BC@AIEAJBw@ABZEED0KV78YZmZmPwQDmpkZPwsAAIA/HAMAAHBCBA==",
    "cellular_caves": "FgIcCS4AAQ@BklCQs@BlRBDNzMw9G@AIMAgAw@ADgC@BCiQIzczMPgkJ@BPkIQH4XrPhjNzEw/DBIkCM3MzD4JCQ@ADBCCAE@BQzczMvhg@B/JAL/BAAL7FE4PgQKFwkNCQg@CQQQDuB4FPwt7FC4/BAOPwnU8DA==",
}

# Trees whose nodes produce NaN / inf / signed zeros at ordinary positions (found by the offline host-emulation fuzz): they pin the
# minps / maxps operand convention of Min, Max, MaxSmooth ... and the fractals' weighted-strength factor over an infinite source.
NON_FINITE_TREES = {
    "max_of_perlin_and_modulus_by_zero": "GwkICP0AAAAYlU2YvwQCLQL/AAAM",                       # Max( Perlin, Modulus( Perlin, 0 ) ): NaN second operand
    "ridged_over_divide_by_zero": "Dwgs/8U/CRgCGQkCAA0zxUEIuq5OvxCyXKG/FAO+2a0+BA==",         # FractalRidged( Divide( Abs( Checkerboard ), 0 ) ): inf source
    "maxsmooth_with_modulus_by_zero": "FgIdCRYCCiABAAAAAzC6P0ALSg/PPgwCLQIEABSeb78IheUQvxB5Puc+GKI0ML8Dzc3XvxvKYShAHA==",
    "constant_minus_zero_under_rotate": "IQkjEDzgJ0AJAAAAAACAFA==",                        # a -0.0f constant keeps its sign (the oracle shim once lost it)
    "cellular_distance_with_nan_jitter": "LgABAAAACSEJCwiaAwAAEK+Xhb8YAAAAgCABAAAAKAIAAAAwAwAAAAIGAND2HEIIMAAAABA8wpk+GNhfST8ECgYAAAAAAAgjAAAAEJIQhz4YEaYIPwQTlqvGPRQ=",
    "cellular_distance_with_nan_size_jitter": "FwkWAgYIdvz//xgAAMBADAILAGONZUEYerruPiAFAAAAKAQAAAACCQBVin1BCCYAAAAQl726Phh2yiU/BBIJAAAAAAAILAAAABCkNKg+GEOjQz8U",
    "max_axis_distance_to_a_nan_point": "KAACAAAAEQIAt63SwQhlJ4m/BAUABAAAAAItAgcAORAfQgjw/f//EAAAwEAYjWjvvwwT4PORQBoFAAQAAAACCwBvEoM6CAQBAAAQ9e/ovxjzQp8/IAIAAAAoAgAAADAAAAAAOP7///8D516dPwtBnTo/E28SgzoECgkAICnpQQgQAwAAEC5w278EEgUAAgAAABsx6bm/IwAAgD8EG9KzhD8iCAAAAACACCUAAAAQ2q3FPhjMRTc/FAouAAAAAAAJLgkKAKwRH0EQ/3e2vxhvEoM6KAIAAAADEtIEQAvCY3w/EyW3iz4c",
    "nested_blends_with_signed_zeros": "KAACAAAAER8CFQkHAJzo+0EIFAAAABCtlkI+GNOPob8EA5a61j4EC8IoTz8EAAQCLQIKGODqZ78gAQAAACgAAAAAAgkAyBgQQQgSAAAAEO2s1z4Y1GczPwQTpzEgPgQKKgkBACf8//8I3hmEPxAYotE+BAPCTD0+DAtF6VG9EzV7Kz4E",
}

# Trees that reach Convert<int> with a value outside [-2^31, 2^31) on FAR_GRID_4D (found by the offline fuzz on a grid far from the
# origin): x86 returns the "integer indefinite" there, the product saturates like ARM / the GPU's own conversion; the C port's
# fn2oSetConvertMode( 1 ) is the oracle for that regime. (seed, encoded tree)
FAR_GRID_4D = dict(off=(-91234.5, 70001.25, -333.0, 1e6), n=(4, 5, 3, 3), step=(37.7, 53.1, 0.5, 1024.0))
CONVERT_OVERFLOW_TREES = [
    (114, 'EAg3AAAAEH0ehz8gGCwxPzABAAAACSIJCAAGHvtBEBrfmr4EChYCBQN9VdQ/E5yijD8bu/fNPyOMEzBABAt6274/BBN/Gas/DA=='),
    (47, 'IgkpAAEAAAAJCwjUAgAAEHL5OL4gBAAAACgAAAAAMAIAAAA4AQAAAAIGACQEFUIIMQAAABBJtQQ+GGf7LD8EC/OJvj4TUH6qPQQDQpYSPwoJCMYAAAAEE8QuAr8bOf52PwQCBQL/AgALgrjoPhIuAAEAAAAJAggb9Ri/EO2I+78MI+0y9j8ECikAAAAAAAkkALU22T8InjuKPxiZmxJACQUABQAAAAthPAdAGxQvoT8MA5PKnT8LB7smPxNlyi2/BBIsAAMAAAAJBQAAAAAAAgcA2NEwQQhh////EPNfsL4YZAZ8vgQLZ+49QBuwfVtAI+OHnD8U'),
    (5, 'IgkUAAMAAAAIEHIeQAkRAG1mVkII1gEAADABAAAACQIA4ywgQghoecM/DAIIACLgeUEIEQAAABDKPuA9GDTCdz4MAwFvhEATsx4VQBoVCQ8AAgAAAAgeCwNACQQAZOW7PhAYKum9GOvbGz8blFmbPwQCCAAbzwBBCCsAAAAQmDPdPRhuRAI/BAshg0Y+FA=='),
    (100, 'JQkFAgwAyxD2QQiT/f//EAQAAAAJCABZzQ5CCKEAAAAQnjz/PxjPHeM/BAshFSE/BBIQAIyhvUEIqf///xAwYiY/GG7avT8JCgA0fQVCGIHphT8gBQAAACgCAAAAC1RUFT8TFEi5PgQCBQO/iIc/CwRyQkATwPYKvxvqBeG/IwtoF0AMG3a4lD8M'),
]



def out_of_range_enum_trees():
    """Enum members written with values outside their range (every setter and the decoder accept any int): the reference's
    switch statements decide -- `default:` labels, fall-throughs, or no case at all. name -> encoded tree"""
    from tests.encode import N
    simplex = lambda: N("Simplex", {"Feature Scale": 23.0})
    return {
        "cellular_distance_function_-1": N("CellularDistance", {"Distance Function": -1, "Feature Scale": 11.0}).encode(),    # EuclideanSquared without the final sqrt
        "cellular_distance_function_9_return_7": N("CellularDistance", {"Distance Function": 9, "Return Type": 7, "Distance Index 1": 2}).encode(),
        "cellular_value_function_6": N("CellularValue", {"Distance Function": 6, "Value Index": 1}).encode(),
        "cellular_lookup_function_-3": N("CellularLookup", {"Distance Function": -3}, [simplex()]).encode(),
        "distance_to_point_function_8": N("DistanceToPoint", {"Distance Function": 8}, [], {("Point", 0): 3.5}).encode(),   # Euclidean with sqrt
        "rotate_plane_type_-1": N("DomainRotatePlane", {"Rotation Type": -1}, [simplex()]).encode(),                         # coordinates pass unchanged
        "rotate_plane_type_2": N("DomainRotatePlane", {"Rotation Type": 2}, [simplex()]).encode(),
        "fade_interpolation_5": N("Fade", {"Interpolation": 5}, [simplex(), N("Perlin")], {"Fade": N("Value", {"Feature Scale": 9.0})}).encode(),
        "warp_simplex_scheme_4": N("DomainWarpSimplex", {"Vectorization Scheme": 4}, [simplex()]).encode(),
        "remove_dimension_6": N("RemoveDimension", {"Remove Dimension": 6}, [simplex()]).encode(),
        "pow_int_-2": N("PowInt", {"Pow": -2}, [simplex()]).encode(),
    }


# checksums of the survey's prototype strict oracle (scalar 4-lane, libm) -- an independent
# build of the same semantics. call = (kind, args): grid3d 64^3 step 4, grid2d 512^2 step 1, grid4d 32,32,16,16 step 4; seed 1337
SURVEY_SMOKE = {
    "fbm_simplex_readme": ("grid2d", "15866cbf", -1.597328, 1.550798),
    "c1_simplex": ("grid3d", "9e65503f", -0.996678, 0.999883),
    "c2_fbm5_perlin": ("grid2d", "7cda0d2b", -1.058768, 1.075703),
    "c3a_cellular_distance": ("grid3d", "23156f56", -0.999909, -0.135814),
    "c3b_cellular_lookup_white": ("grid3d", "e3ae3be6", -0.983556, 0.968467),
    "c5_supersimplex": ("grid4d", "d9cee143", -0.790859, 0.796379),
    "c4_terrain": ("grid3d", "be4f2cf0", 0.166667, 0.928188),
    "cellular_caves": ("grid3d", "06837b6b", -2.118493, 0.076068),
    "mountain_terrain": ("grid3d", "40043e0d", -0.372797, 2.460488),
}


def survey_call(kind):
    if kind == "grid2d":
        return dict(kind="grid2d", off=(0.0, 0.0), n=(512, 512), step=(1.0, 1.0), seed=1337)
    if kind == "grid3d":
        return dict(kind="grid3d", off=(0.0, 0.0, 0.0), n=(64, 64, 64), step=(4.0, 4.0, 4.0), seed=1337)
    return dict(kind="grid4d", off=(0.0, 0.0, 0.0, 0.0), n=(32, 32, 16, 16), step=(4.0, 4.0, 4.0, 4.0), seed=1337)


def golden_positions(n, dims, seed=1337):
    rng = np.random.default_rng(seed)
    return [rng.uniform(-1000.0, 1000.0, n).astype(np.float32) for _ in range(dims)]


GOLDEN_CASES = {
    "c1_simplex_grid3d": dict(tree=TREES["c1_simplex"], kind="grid3d", off=(-5.3, 2.1, 0.7), n=(28, 24, 20), step=(3.7, 3.7, 3.7), seed=1337),
    "c2_fbm5_perlin_grid2d": dict(tree=TREES["c2_fbm5_perlin"], kind="grid2d", off=(-17.0, 4.5), n=(96, 80), step=(0.25, 0.25), seed=1337),
    "c2_fbm5_perlin_tileable": dict(tree=TREES["c2_fbm5_perlin"], kind="tileable2d", n=(72, 56), step=(1.0, 1.0), seed=1337),
    "c3a_cellular_distance_grid3d": dict(tree=TREES["c3a_cellular_distance"], kind="grid3d", off=(0.0, 0.0, 0.0), n=(24, 24, 20), step=(5.1, 5.1, 5.1), seed=1337),
    "c3b_cellular_lookup_white_grid3d": dict(tree=TREES["c3b_cellular_lookup_white"], kind="grid3d", off=(0.0, 0.0, 0.0), n=(24, 24, 20), step=(5.1, 5.1, 5.1), seed=1337),
    "c3c_cellular_lookup_gradient_grid3d": dict(tree=TREES["c3c_cellular_lookup_gradient"], kind="grid3d", off=(0.0, 0.0, 0.0), n=(24, 24, 20), step=(5.1, 5.1, 5.1), seed=1337),
    "c4_terrain_grid3d": dict(tree=TREES["c4_terrain"], kind="grid3d", off=(0.0, 0.0, 0.0), n=(28, 24, 20), step=(6.3, 6.3, 6.3), seed=1337),
    "c5_supersimplex_pos4d": dict(tree=TREES["c5_supersimplex"], kind="pos4d", count=4099, off=(0.0, 0.0, 0.0, 0.0), seed=1337),
    "cellular_caves_grid3d": dict(tree=DEMO_TREES["cellular_caves"], kind="grid3d", off=(1.0, -2.0, 3.0), n=(20, 20, 16), step=(2.9, 2.9, 2.9), seed=-7),
    "mountain_terrain_grid3d": dict(tree=DEMO_TREES["mountain_terrain"], kind="grid3d", off=(-40.0, 10.0, 25.0), n=(24, 20, 16), step=(9.7, 9.7, 9.7), seed=1337),
    "fbm_simplex_grid4d": dict(tree=TREES["fbm_simplex_readme"], kind="grid4d", off=(0.5, 0.25, -0.75, 9.0), n=(12, 10, 8, 6), step=(7.0, 7.0, 7.0, 7.0), seed=0x7fffffff),
}


def run_case(node, case):
    """node: anything with grid2d/grid3d/grid4d/tileable2d/positions methods returning (out, minmax) -- tests.helpers.RefNode,
    tests.helpers.PortNode or the product adapter in tests.helpers.ProductNode."""
    k = case["kind"]
    if k == "grid2d":
        return node.grid2d(*case["off"], *case["n"], *case["step"], case["seed"])
    if k == "grid3d":
        return node.grid3d(*case["off"], *case["n"], *case["step"], case["seed"])
    if k == "grid4d":
        return node.grid4d(*case["off"], *case["n"], *case["step"], case["seed"])
    if k == "tileable2d":
        return node.tileable2d(*case["n"], *case["step"], case["seed"])
    dims = int(k[3])
    return node.positions(golden_positions(case["count"], dims), case["off"][:dims], case["seed"])
