import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu through gpurun)")
    config.addinivalue_line("markers", "timeout(seconds): pytest-timeout's marker, registered here too so that a box without the plugin only ignores it")


def _has_gpu() -> bool:
    try:
        import ctypes
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        cu.cuDeviceGetCount(ctypes.byref(n))
        return n.value > 0
    except OSError:
        return False


HAS_GPU = _has_gpu()

# FCP_TEST_EMU=1: run the `-m gpu` suite against tests/emu/libfcp_emu.so, the product's csrc/*.cu re-compiled for a CPU
# emulation of the CUDA execution model (tests/emu/cuda_runtime.h).  Test infrastructure for containers without a GPU:
# the product itself (freecappuccino-dev_b200/lib.py) has no such switch and no CPU path.
import emu_hook  # noqa: E402
EMU = emu_hook.wanted() and not HAS_GPU
if EMU:
    emu_hook.activate()
    HAS_GPU = True


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no GPU in this container (GPU tests run through gpurun / the driver)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    from oracle import orc_py
    orc_py.lib()
    return orc_py


@pytest.fixture(scope="session")
def fcp():
    """The product package; the CUDA library must already be built (no silent fallback)."""
    import fcb200
    from fcb200 import lib
    lib.lib()
    return fcb200
