import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _have_gpu():
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library, initialised on device 0.  Fails loudly if it cannot be loaded: a GPU test
    must never pass on a fallback."""
    import hpgmg_b200.api as api
    assert _have_gpu(), "no NVIDIA device visible: -m gpu tests need the B200 box"
    api.init(0)
    L = api.lib()
    assert L.hpgmg_b200_backend() == b"cuda-sm_100a"
    L.hpgmg_b200_set_layout_only(0)
    L.hpgmg_b200_set_verbose(0)
    return L


@pytest.fixture(scope="session")
def layout_lib():
    """The same library in layout-only mode (host data model, no device) for the CPU suite."""
    import hpgmg_b200.api as api
    L = api.lib()
    L.hpgmg_b200_set_layout_only(1)
    L.hpgmg_b200_set_verbose(0)
    L.hpgmg_b200_set_agglomeration(0)          # the reference's own rank_of_box: these tests diff the lists against it
    return L
