"""robustrmab_b200 -- B200 (sm_100a) implementation of the RobustRMAB hot path.

Layers (bottom up):
  csrc/ + include/rmab.h   hand-written CUDA kernels behind a C ABI (librmab_b200.so)
  _lib.py                  ctypes binding of that ABI
  ops.py                   torch-tensor front end (device memory + stream plumbing only)
There is no CPU fallback anywhere in this package: without the built library, or with CPU tensors,
every entry point raises.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops"]


def __getattr__(name):
    if name == "ops":
        import importlib
        return importlib.import_module(".ops", __name__)
    raise AttributeError(name)
