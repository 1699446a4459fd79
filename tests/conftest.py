import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """libsdfk.so is built in-tree (git-ignored); from a clean checkout build it once, like __graft_entry__.build() does.
    (A no-op when the library is newer than its sources.  The product itself never builds or falls back on its own.)"""
    from sdfkit_b200 import build
    build.build()
