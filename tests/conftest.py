import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    import tenet_jl_b200 as tb
    return tb.load_library()


@pytest.fixture(scope="session")
def ctx(built_lib):
    import tenet_jl_b200 as tb
    return tb.default_context()
