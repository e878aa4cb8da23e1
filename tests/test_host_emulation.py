"""The lowering, checked on the CPU. The CUDA source that lower.cpp generates for a tree (plus the device node library it
includes) is compiled with the HOST compiler over tests/host_emulation/cuda_host_shim.h and run "thread" by "thread"; the
values must equal the strict oracle's bit for bit. This exercises the emitter and the scalar and packed instantiations of
the node library for random trees without a GPU. It is a test harness around product-generated code, not a product path:
nothing under fastnoise2_b200/ can evaluate noise on the CPU.

    python tests/test_host_emulation.py 2000        # offline fuzz with that many random trees per flavour
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


def run_on_host(fn, text, dims, seed, workdir, tag):
    """Lower `text` for a dims-D grid, compile the generated CUDA with g++ over the stand-in, run it, return the values."""
    _BUILDS[0] += 1
    tag = "%s_%d" % (tag, _BUILDS[0])  # dlopen() hands back an already loaded library for a path it has seen
    source = fn.Generator(text).kernel_source(dims)
    ppt = int(source.split(" point(s) per thread")[0].rsplit(" ", 1)[1])
    path = os.path.join(workdir, "tree_%s.cpp" % tag)
    with open(path, "w") as f:
        f.write('#include "cuda_host_shim.h"\n' + source + '\n#include "driver.inl"\n')
    lib_path = os.path.join(workdir, "tree_%s.so" % tag)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fno-fast-math", "-fwrapv", "-w", "-fPIC", "-shared",
                           "-I" + os.path.join(ROOT, "tests", "host_emulation"), "-I" + os.path.join(ROOT, "fastnoise2_b200", "csrc", "device"), path, "-o", lib_path])
    lib = ctypes.CDLL(lib_path)
    g = GRIDS[dims]
    out = np.zeros(int(np.prod(g["n"])), np.float32)
    lib.fn2_host_grid(out.ctypes.data_as(ctypes.c_void_p), dims, (ctypes.c_float * dims)(*g["off"]), (ctypes.c_int * dims)(*g["n"]),
                      (ctypes.c_float * dims)(*g["step"]), seed, ppt)
    return out


def oracle_values(node, dims, seed):
    g = GRIDS[dims]
    call = {2: node.grid2d, 3: node.grid3d, 4: node.grid4d}[dims]
    return call(*g["off"], *g["n"], *g["step"], seed)[0]


def fuzz(fn, ref_lib, count, seed, workdir, label):
    fuzzer = TreeFuzzer(seed)
    failures = []
    for index in range(count):
        text = fuzzer.tree(depth=3).encode()
        for dims in (2, 3, 4):
            got = run_on_host(fn, text, dims, index, workdir, "%s_%d" % (label, dims))
            want = oracle_values(helpers.RefNode(ref_lib, text), dims, index)
            if not same(got, want):
                failures.append((label, index, dims, text, helpers.compare(got, want)))
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
    bad = []
    for variant, mode, label in (("strict_avx512", 0, "tables"), ("strict_avx2", 2, "generic")):
        product.set_math_mode(mode)
        bad += fuzz(product, helpers.load_ref(variant), total, 9000 + mode, tempfile.mkdtemp(), label)
        print(label, "done,", len(bad), "mismatching (tree, dimension) pairs so far", flush=True)
    for item in bad[:20]:
        print(item)
