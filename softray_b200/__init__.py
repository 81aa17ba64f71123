"""softray_b200 -- B200-native (sm_100a) drop-in for SoftRay's per-pixel raytrace hot path.

The product is libsoftray_cuda.so (CUDA kernels + C ABI, include/softray_cuda.h).  This package is
the thin host side above it: ctypes bindings (`abi`, `lib`), plain data containers (`scene`) and a
Python mirror of the reference's Engine3D host surface (`engine3d`).
There is no CPU fallback: without the built extension or without a CUDA device every render call
raises.
"""
from . import abi  # noqa: F401
from .scene import FrameParams, InstanceData, MeshData, SceneDescHolder, SphereData  # noqa: F401

__all__ = ["abi", "FrameParams", "InstanceData", "MeshData", "SceneDescHolder", "SphereData"]
