import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hmrm():
    """The product package (ctypes over libhmrm.so); built in-tree if stale."""
    import __graft_entry__ as entry

    entry.build_product()
    import hmrm_pkg

    return hmrm_pkg.load()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def renderer(hmrm):
    r = hmrm.Renderer(0)
    yield r
    r.close()
