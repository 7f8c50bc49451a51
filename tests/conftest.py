import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    oracle.set_threads(min(8, os.cpu_count() or 1))
    return oracle


@pytest.fixture(scope="session")
def fcb_lib():
    """The CUDA library; built in-tree if stale. GPU tests must fail (not skip) if it cannot load."""
    from simple_image_compression_network_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.lib()


@pytest.fixture
def exp_build(fcb_lib):
    """Cross-check tests only: handles created after `exp_build()` is called come from the EXPERIMENT build of the same sources
    (tools/libfinnconv_exp.so: -DFCB_EXPERIMENT adds the environment switches that bend plans, the first-generation kernel and the
    cross-check instantiations).  The product library has none of them.  The default is restored when the test ends."""
    from simple_image_compression_network_b200 import _lib, build
    state = {}

    def switch():
        if "L" not in state:
            if not os.path.exists(_lib.EXP_LIB_PATH):
                build.build(exp=True)
            state["L"] = _lib.load(_lib.EXP_LIB_PATH)
        _lib.set_default(state["L"])
        return state["L"]

    yield switch
    _lib.set_default(None)
