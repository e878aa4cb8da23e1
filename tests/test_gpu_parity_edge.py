"""Edge regimes of the parity contract, through the C ABI (needs a B200: pytest -m gpu). Collected after test_gpu_parity.py.

Values that are not finite, enum members that are out of range and cell coordinates that do not fit an int32: the regimes the
offline host-emulation fuzz (tests/test_host_emulation.py) found differences in, now pinned. The same
cases run on the CPU against the generated code in tests/test_host_emulation.py.
"""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(fn):
    assert fn.device_available(), "these tests need a CUDA device: " + fn.last_error()
    return fn


def test_non_finite_values_inside_a_tree_follow_the_x86_oracle(gpu, ref):
    """Divide / Modulus by zero feed NaN and inf into Min, Max, MaxSmooth and a fractal: the oracle's minps / maxps operand
    convention and its NaN weighted-strength factor decide the output (tests/cases.py NON_FINITE_TREES)."""
    from tests.cases import NON_FINITE_TREES
    from tests.fuzz import same
    failures = []
    for name, text in NON_FINITE_TREES.items():
        a, b = helpers.ProductNode(gpu, text), helpers.RefNode(ref, text)
        for dim, got, want in ((2, a.grid2d(-3.3, 2.1, 48, 40, 1.9, 1.7, 146)[0], b.grid2d(-3.3, 2.1, 48, 40, 1.9, 1.7, 146)[0]),
                               (3, a.grid3d(-3.3, 2.1, 0.7, 16, 14, 12, 1.9, 1.7, 2.3, 146)[0], b.grid3d(-3.3, 2.1, 0.7, 16, 14, 12, 1.9, 1.7, 2.3, 146)[0])):
            if not same(got, want):
                failures.append((name, dim, helpers.compare(got, want)))
    assert not failures, failures


def test_out_of_range_enum_values_follow_the_reference_switches(gpu, ref):
    """Enum members written outside their range (setters and decoder accept any int) take the reference's `default:` label,
    fall-through or no-case path (tests/cases.py out_of_range_enum_trees)."""
    from tests.cases import out_of_range_enum_trees
    from tests.fuzz import same
    failures = []
    for name, text in out_of_range_enum_trees().items():
        a, b = helpers.ProductNode(gpu, text), helpers.RefNode(ref, text)
        for dim, got, want in ((2, a.grid2d(-3.3, 2.1, 48, 40, 1.9, 1.7, 31)[0], b.grid2d(-3.3, 2.1, 48, 40, 1.9, 1.7, 31)[0]),
                               (3, a.grid3d(-3.3, 2.1, 0.7, 16, 14, 12, 1.9, 1.7, 2.3, 31)[0], b.grid3d(-3.3, 2.1, 0.7, 16, 14, 12, 1.9, 1.7, 2.3, 31)[0]),
                               (4, a.grid4d(-3.3, 2.1, 0.7, -1.9, 8, 7, 6, 5, 1.9, 1.7, 2.3, 1.1, 31)[0], b.grid4d(-3.3, 2.1, 0.7, -1.9, 8, 7, 6, 5, 1.9, 1.7, 2.3, 1.1, 31)[0])):
            if not same(got, want):
                failures.append((name, dim, helpers.compare(got, want)))
    assert not failures, failures


def test_convert_overflow_regime_matches_the_saturating_port(gpu, port):
    """Positions whose cell coordinates leave the int32 range: x86 builds of the reference produce the "integer indefinite",
    this implementation saturates (like ARM builds); the C restatement switched to saturating conversion is the oracle."""
    from tests.cases import CONVERT_OVERFLOW_TREES, FAR_GRID_4D
    from tests.fuzz import same
    g = FAR_GRID_4D
    port.fn2oSetConvertMode(1)
    try:
        for seed, text in CONVERT_OVERFLOW_TREES:
            got = helpers.ProductNode(gpu, text).grid4d(*g["off"], *g["n"], *g["step"], seed)[0]
            want = helpers.PortNode(port, text).grid4d(*g["off"], *g["n"], *g["step"], seed)[0]
            assert same(got, want), (text, helpers.compare(got, want))
    finally:
        port.fn2oSetConvertMode(0)


def test_short_lived_caller_threads_do_not_leak_device_memory(gpu):
    """Each calling thread owns streams and staging buffers per device; they are released when the thread exits (a consumer that
    spawns worker threads per chunk -- the Node Editor's mesh preview does -- must not lose 128 MiB of HBM per thread)."""
    import threading
    import torch
    g = gpu.Generator(gpu.TREES["c1_simplex"])
    n = 96

    def work(results, index):
        out = np.empty(n ** 3, np.float32)
        results[index] = g.grid3d(out, float(index), 0.0, 0.0, n, n, n, 1.0, 1.0, 1.0, 1337)

    def churn(count):
        results = [None] * count
        for i in range(count):
            t = threading.Thread(target=work, args=(results, i))
            t.start()
            t.join()
        assert all(r is not None and r[0] < r[1] for r in results)

    churn(3)
    torch.cuda.synchronize()
    free_before, _ = torch.cuda.mem_get_info()
    churn(24)
    torch.cuda.synchronize()
    free_after, _ = torch.cuda.mem_get_info()
    assert free_before - free_after < 16 << 20, (free_before, free_after)


def test_grid_limits_are_refused_not_miscomputed(gpu):
    """Index arithmetic is exact below 2^31 rows / points; anything larger is refused with an error (never silently decomposed wrongly)."""
    import torch
    g = gpu.Generator(gpu.TREES["c1_simplex"])
    out = torch.empty(16, dtype=torch.float32, device="cuda")
    rc = gpu.lib.fnb200GenUniformGrid3D(g.handle, out.data_ptr(), 0.0, 0.0, 0.0, 1, 65536, 65536, 1.0, 1.0, 1.0, 1337, 40000, 40001, None, None)
    assert rc == -1 and "2^31" in gpu.last_error()
    mm = np.zeros(2, np.float32)
    gpu.lib.fnGenUniformGrid4D(g.handle, out.data_ptr(), 0.0, 0.0, 0.0, 0.0, 65536, 65536, 65536, 65536, 1.0, 1.0, 1.0, 1.0, 1337, mm.ctypes.data)
    assert mm[0] > mm[1] and "2^31" in gpu.last_error()
