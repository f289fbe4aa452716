"""mcray_tracing_b200 -- B200-native (sm_100a) per-frame hot path of thepochynsons/MCRay-Tracing.

The product is `libmcrt.so` (C ABI: include/mcrt.h; sources under csrc/).  This package is the
thin Python host mirror: `api` (ctypes binding), `assets` (synthetic stand-ins for the meshes the
reference does not ship) and `sweep` (one-process-per-GPU probe-sweep sharding over NCCL).
"""
from . import api, assets  # noqa: F401

__all__ = ["api", "assets"]
