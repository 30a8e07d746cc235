import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session if they are missing."""
    from arendur_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def ctx():
    from arendur_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def cornell_small():
    """C1-like: Cornell box at 96x72, 2x2 spp (host scene + camera + film + sampler + params)."""
    from arendur_b200 import scenes
    return scenes.cornell_scene(96, 72, 2, 2)
