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


# The reference has two gradient FLAVOURS (compile-time ISA branches that round differently); the product implements both
# and each is checked bit for bit against its own build of the reference sources and its own golden vectors.
#   tables  = the AVX512_F permute-table branches  (product default, fnb200SetMathMode(0))
#   generic = the select/xor branches of the other ISA levels (fnb200SetMathMode(2))
FLAVOURS = {
    "tables": dict(ref="strict_avx512", port=1, mode=0, golden="_tables"),
    "generic": dict(ref="strict_avx2", port=0, mode=2, golden=""),
}


@pytest.fixture(scope="module", params=["tables", "generic"])
def flavour(request, fn):
    import types
    f = types.SimpleNamespace(name=request.param, **FLAVOURS[request.param])
    fn.set_math_mode(f.mode)
    yield f
    fn.set_math_mode(0)


@pytest.fixture(scope="module")
def ref(flavour):
    """Strict oracle of the flavour under test: the reference's node sources over oracle/shim (oracle/_ref/libfn2ref_strict_*.so)."""
    from tests import helpers
    lib = helpers.load_ref(flavour.ref)
    if lib is None:
        pytest.skip("oracle/_ref/libfn2ref_%s.so not built or not runnable on this CPU (needs /root/reference: make -C oracle)" % flavour.ref)
    return lib


@pytest.fixture(scope="module")
def port(flavour):
    """The plain-C restatement (oracle/fn2_oracle.c) switched to the flavour under test."""
    from tests import helpers
    lib = helpers.load_port()
    if lib is None:
        pytest.skip("oracle/libfn2oracle.so not built (make -C oracle port)")
    lib.fn2oSetFlavor(flavour.port)
    return lib
