import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run under gpurun on a B200)")


@pytest.fixture(scope="session")
def emu_lib():
    """CPU emulation build of the kernel sources (tests/emu) - checks kernel logic without a GPU."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from ark_plonk_b200._lib import Lib
    lib = Lib(build_emu.build())
    lib.init()
    return lib


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real GPU.  No fallback: a missing build or device is a failure."""
    from ark_plonk_b200 import build as apb_build
    from ark_plonk_b200._lib import get_lib
    apb_build.build()
    lib = get_lib()
    lib.init(0)
    return lib
