"""libmultiviewnative_b200 -- B200-native multi-view Richardson-Lucy deconvolution
behind the reference's C API (psteinb/libmultiviewnative, inc/multiviewnative.h).

The product is ``lib/libmultiviewnative.so`` (hand-written sm_100a CUDA + a C ABI);
this package is the thin Python host layer over it: ``capi`` (ctypes mirror of
the reference interface + the persistent plan handle), ``synthetic`` (seeded
workloads), ``blocks`` (sharding independent blocks over the GPUs of one box).
"""
from .capi import Library, LmvnError, Plan, load  # noqa: F401

__all__ = ["Library", "LmvnError", "Plan", "load"]
