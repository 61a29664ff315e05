"""B200-native viewpoint-gain path of the NBV exploration planner.

``NbvGpu`` wraps the C ABI (include/nbvgpu.h -> csrc/libnbvgpu.so, hand-written sm_100a CUDA);
``RrtGain`` / ``LowResManager`` / ``HighResManager`` mirror the reference's call sites.  There is no
CPU implementation in this package: constructing ``NbvGpu`` without a B200 raises.
"""
from ._native import (LAYER_ESDF, LAYER_TSDF, PACK_PLANAR, PACK_VOXBLOX, NbvGpuError, build)
from .planner import CameraParams, HighResManager, History, LowResManager, NbvGpu, RrtGain, kFree, kOccupied, kUnknown

__all__ = ["NbvGpu", "RrtGain", "LowResManager", "HighResManager", "History", "CameraParams", "NbvGpuError", "build",
           "LAYER_TSDF", "LAYER_ESDF", "PACK_PLANAR", "PACK_VOXBLOX", "kUnknown", "kFree", "kOccupied"]
