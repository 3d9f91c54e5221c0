"""freecappuccino-dev_b200 -- B200-native pressure-velocity coupling hot path of freeCappuccino.

The directory name carries a hyphen, so import it with
``importlib.import_module("freecappuccino-dev_b200")`` (or ``import fcb200`` from the repo root).

Sub-modules
-----------
mesh    host-side mesh layer (geometry module mirror, fixture generators, partitioner)
lib     ctypes binding of the C-ABI shared library ``csrc/libfcp_b200.so`` (include/fcp.h)
host    mirror of the reference's module-level API (csrsolve, grad, laplacian, calcp_simple, exchange ...)

There is no CPU fallback: every compute entry point goes through the CUDA library and raises
``RuntimeError`` when the library is missing or no B200 is visible.
"""
from . import mesh  # noqa: F401

__all__ = ["mesh", "lib", "host"]


def __getattr__(name):
    if name in ("lib", "host", "build"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
