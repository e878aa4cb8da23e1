"""-m gpu smoke tests of the tools around the path: the FastNoiseBenchmark-shaped harness on the product, incl. its "Node Trees" group."""
import os
import subprocess
import sys

import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def test_node_benchmark_product_arm_runs_node_trees_and_a_node(fn):
    if not fn.device_available():
        pytest.skip("no CUDA device")
    tool = os.path.join(helpers.ROOT, "tools", "node_benchmark.py")
    out = subprocess.run([sys.executable, tool, "--filter", "node trees", "--dims", "2", "3", "4", "--positions", "262144", "--iterations", "3"],
                         capture_output=True, text=True, cwd=helpers.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [line for line in out.stdout.splitlines() if line.strip()]
    assert len(lines) == 6 and all("/B200_SM100A/Node Trees/" in line and "M items/s" in line for line in lines)
    assert sum("Mountain Terrain" in line for line in lines) == 3 and sum("Cellular Caves" in line for line in lines) == 3
    out = subprocess.run([sys.executable, tool, "--filter", "Perlin", "--dims", "3", "--positions", "262144", "--iterations", "3"],
                         capture_output=True, text=True, cwd=helpers.ROOT)
    assert out.returncode == 0 and "3D/B200_SM100A/Coherent Noise/Perlin" in out.stdout, out.stderr[-2000:]
