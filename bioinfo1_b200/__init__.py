"""bioinfo1_b200 -- B200-native batched pairwise alignment + minimizer extraction.

The product is the native library (csrc/ -> libb200map.so, C ABI in include/b200map.h) and the
C++ drop-in wrappers (libteam_b200.so). This Python package is only the harness around it:
`build` compiles the native code in-tree, `capi` binds the C ABI with ctypes for tests and
bench.py. There is no Python or CPU implementation of the hot path in here.
"""
from . import build  # noqa: F401

__all__ = ["build", "capi"]
