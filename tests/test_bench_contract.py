"""bench.py contract, as far as it can be checked without a GPU: the reference arm (the reference's node code on the host
cores) runs, prints exactly one JSON line and carries the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    if helpers.load_ref("relaxed_avx2") is None and helpers.load_ref("relaxed_avx512") is None:
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "Mpoints/s" and line["higher_is_better"] is True and line["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["steps"] == 2 and line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 times the CPU reference; the other ranks print nothing and exit 0."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout, out.stderr[-500:])


def test_node_benchmark_reference_arm_runs():
    """tools/node_benchmark.py (the FastNoiseBenchmark-shaped harness) on the CPU reference arm, two nodes, reference naming."""
    if helpers.load_ref("relaxed_avx2") is None and helpers.load_ref("relaxed_avx512") is None:
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "node_benchmark.py"), "--impl", "reference", "--filter", "Perlin", "--dims", "2", "3",
                          "--positions", "1024", "--iterations", "3", "--cpu-seconds", "0.05"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.strip()]
    assert len(lines) == 2 and lines[0].startswith("2D/") and lines[1].startswith("3D/")
    assert all("/Coherent Noise/Perlin" in line and "M items/s" in line for line in lines)


def test_node_benchmark_reference_arm_has_the_node_trees_group():
    """The "Node Trees" group of the reference harness (FastNoiseBenchmark.cpp:152-166): both demo trees, every dimensionality."""
    if helpers.load_ref("relaxed_avx2") is None and helpers.load_ref("relaxed_avx512") is None:
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "node_benchmark.py"), "--impl", "reference", "--filter", "node trees", "--dims", "2", "3", "4",
                          "--positions", "1024", "--iterations", "2", "--cpu-seconds", "0.05"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.strip()]
    assert len(lines) == 6 and all("/Node Trees/" in line for line in lines)
    assert [("Mountain Terrain" in line) for line in lines] == [True] * 3 + [False] * 3


def test_both_arms_describe_the_same_config():
    """The driver compares the `config` of the two arms' lines: it comes from one function, fed the same arguments."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    for workload, world, scaling in (("c4", 1, "strong"), ("c4", 8, "strong"), ("c5", 2, "weak")):
        args = argparse.Namespace(workload=workload, scaling=scaling)
        a = bench.config_of(args, bench.WORKLOADS[workload], world)
        b = bench.config_of(argparse.Namespace(workload=workload, scaling=scaling), bench.WORKLOADS[workload], world)
        assert a == b and a["encoded_tree"] and a["points"] > 0 and "exceeds L2" in a["l2"]
    assert bench.slab_range(1024, 7, 8) == (896, 1024) and bench.slab_range(10, 0, 3) == (0, 4)
