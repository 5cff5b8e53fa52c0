"""dynamicfusion_b200 -- B200-native (sm_100a) implementation of the DynamicFusion per-frame hot path.

The product is libdfusion.so (hand-written CUDA behind the C ABI of include/dfusion.h).  This package holds
the build script, the ctypes binding and the Python host-side mirror of the reference's KinFu / TsdfVolume /
WarpField interface used by tests and bench.py.  There is no CPU fallback.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
