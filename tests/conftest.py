import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure libb200rt.so and the oracle's C restatement exist (compiles them if needed)."""
    import __graft_entry__ as g
    from libyafaray_b200 import rt
    if not os.path.exists(rt.LIB_PATH):
        g.build()
    from oracle import kdo
    kdo.build_library()
    return True
