import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


@pytest.fixture(scope='session')
def built():
    """Builds (or reuses) every native artefact once per session."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope='session')
def ref(built):
    from oracle import pyoracle
    return pyoracle.Ref()


@pytest.fixture(scope='session')
def emu(built):
    from tests.common import HostEmu
    return HostEmu()


@pytest.fixture(scope='session')
def emu_big(built):
    """DiscEngine under the host policy compiled for the large-grid build (Nx up to 16384)."""
    from tests.common import HostEmu
    return HostEmu(big=True)
