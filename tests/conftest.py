import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (intrinsictimescales.jl_b200/) under the module name itjl_b200."""
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def ctx(pkg):
    """A GPU context; GPU tests fail loudly (not skip) when the library cannot run."""
    return pkg.default_context(0)
