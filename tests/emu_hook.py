"""TEST INFRASTRUCTURE ONLY.  FCP_TEST_EMU=1 on a machine without a GPU points the ctypes binding at
tests/emu/libfcp_emu.so -- the product's csrc/*.cu re-compiled against a CPU emulation of the CUDA execution model
(tests/emu/cuda_runtime.h) -- so that the `-m gpu` parity suite and the multi-rank worker can exercise the kernels' logic
here.  The product (freecappuccino-dev_b200/lib.py) has no such switch: it only ever loads csrc/libfcp_b200.so."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def wanted() -> bool:
    return os.environ.get("FCP_TEST_EMU", "") == "1"


def activate() -> str:
    """Build (if stale) and select the emulation library; returns its path."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import build as emu_build
    import fcb200  # noqa: F401
    from fcb200 import lib
    lib.SO_PATH = emu_build.build()
    lib._LIB = None
    return lib.SO_PATH
