"""cuda_flow3d_b200 -- B200-native (sm_100a) dense 3D variational optical flow.

A from-scratch replacement for the hot path of axruff/cuda-flow3d
(OpticalFlowE::ComputeFlow and the six CUDA operations under it).  The product is
libflow3d_b200.so (hand-written CUDA kernels behind the C ABI in include/flow3d_c.h) plus the C++
host classes in host/ that keep the reference's API; this Python package is the thin ctypes mirror
used by the tests and bench.py.  No CPU fallback exists.
"""
from ._lib import Flow3DError, Params, load, require_device, LIB_PATH  # noqa: F401
from .api import (DEFAULTS, Data3D, DataSize4, DeviceVolume, OperationParameters, OpticalFlowE,  # noqa: F401
                  level_schedule, ops)
