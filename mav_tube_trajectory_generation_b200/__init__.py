"""B200-native batched min-snap trajectory solver (hot path of
NilsFunk/mav_tube_trajectory_generation).

The product is ``libmtg_cuda.so`` (hand-written sm_100a kernels behind the C ABI
of ``include/mtg_cuda.h``) plus the C++ class shim under
``include/mav_tube_trajectory_generation``. This Python package only builds and
binds the library for tests and ``bench.py``.
"""
from . import _build, sweep  # noqa: F401
from .capi import Context, MtgError, get_tables, load  # noqa: F401

__all__ = ["Context", "MtgError", "get_tables", "load"]
