import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fn():
    """The product library binding. Built by __graft_entry__.build(); import fails loudly if it is missing."""
    import fastnoise2_b200
    return fastnoise2_b200


@pytest.fixture(scope="session")
def ref():
    """Strict oracle: the reference's node sources over oracle/shim (oracle/_ref/libfn2ref_strict_avx2.so)."""
    from tests import helpers
    lib = helpers.load_ref("strict_avx2")
    if lib is None:
        pytest.skip("oracle/_ref/libfn2ref_strict_avx2.so not built (needs /root/reference: make -C oracle)")
    return lib
