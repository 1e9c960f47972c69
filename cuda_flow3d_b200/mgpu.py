"""ctypes mirror of the z-sharded multi-GPU solver (libflow3d_b200_mgpu.so, include/flow3d_mgpu_c.h).

The solver itself is C++ (csrc/sharded_solver.cu: the C ABI's *_slab stage calls + NCCL send/recv over
NVLink); this module only binds it, for bench.py (one process per GPU under torchrun) and the tests.
Replaces the role of the reference's out-of-core slab driver
(src/cuda_operations/partial_data/cuda_operation_solve_p.cpp:358-417, src/optical_flow/optical_flow_p.cpp:58-323).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import Params, check
from .api import make_params

MGPU_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libflow3d_b200_mgpu.so")
ID_BYTES = 128
PHASES = ["prolongation", "flow_ghost_exchange", "level_frames", "warp_derivs", "solver", "halo_exchange", "update",
          "median", "blur"]
STATS = ["sharded_levels", "replicated_levels", "exchanges", "exchange_bytes_sent", "frame_gathers", "voxel_sweeps",
         "phi_voxels", "peak_device_bytes"]

_vp = C.c_void_p
_szp = C.POINTER(C.c_size_t)
SIGNATURES = {
    "flow3d_mgpu_unique_id": (C.c_int, [_vp]),
    "flow3d_sharded_own_range": (None, [C.c_size_t, C.c_int, C.c_int, _szp, _szp]),
    "flow3d_sharded_input_planes": (None, [C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_size_t, _szp, _szp]),
    "flow3d_sharded_create": (C.c_int, [C.c_size_t] * 3 + [C.c_int] * 3 + [_vp, C.POINTER(_vp)]),
    "flow3d_sharded_destroy": (C.c_int, [_vp]),
    "flow3d_sharded_compute": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(Params),
                                         C.c_size_t, _vp, _vp, _vp, C.c_size_t, _szp, _szp, _vp]),
    "flow3d_sharded_output_planes": (C.c_int, [_vp, C.POINTER(Params), _szp, _szp]),
    "flow3d_sharded_plan_check": (C.c_int, [C.c_size_t] * 3 + [C.c_int, C.POINTER(Params), C.c_size_t, C.c_size_t, C.c_size_t,
                                            C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "flow3d_sharded_frame_ghost": (C.c_size_t, [C.c_size_t] * 3 + [C.c_int, C.POINTER(Params), C.c_size_t, C.c_size_t]),
    "flow3d_sharded_set_thresholds": (C.c_int, [_vp, C.c_size_t, C.c_size_t]),
    "flow3d_sharded_set_profiling": (C.c_int, [_vp, C.c_int]),
    "flow3d_sharded_phase_ms": (C.c_int, [_vp, C.c_float * 9]),
    "flow3d_sharded_stats": (C.c_int, [_vp, C.c_double * 8]),
    "flow3d_sharded_tune": (C.c_int, [_vp, C.POINTER(Params)]),
    "flow3d_mgpu_compute_host": (C.c_int, [C.c_size_t] * 3 + [C.c_int, C.POINTER(C.c_int), _vp, _vp, C.POINTER(Params),
                                           _vp, _vp, _vp, C.POINTER(C.c_float), C.c_int]),
}

_mlib = None


def load():
    """Load libflow3d_b200_mgpu.so (after the single-GPU library it is linked against).  Loud failure if it
    was not built: there is no fallback."""
    global _mlib
    if _mlib is None:
        _lib.load()
        if not os.path.exists(MGPU_LIB_PATH):
            raise ImportError("%s not found: build it with `make` (needs NCCL)" % MGPU_LIB_PATH)
        lib = C.CDLL(MGPU_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _mlib = lib
    return _mlib


def unique_id():
    buf = C.create_string_buffer(ID_BYTES)
    check(load().flow3d_mgpu_unique_id(buf), "flow3d_mgpu_unique_id")
    return buf.raw


def own_range(d, rank, world):
    a, b = C.c_size_t(), C.c_size_t()
    load().flow3d_sharded_own_range(d, rank, world, C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def input_planes(depth, rank, world, sigma, frame_ghost):
    a, b = C.c_size_t(), C.c_size_t()
    load().flow3d_sharded_input_planes(depth, rank, world, sigma, frame_ghost, C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def plan_check(W, H, D, world, params=None, min_planes=12, min_voxels=1 << 18, frame_ghost=32):
    """host-only dry run of the partition arithmetic; returns (ok, bad_level, bad_rank)"""
    p = make_params(params)
    bl, br = C.c_int(-1), C.c_int(-1)
    rc = load().flow3d_sharded_plan_check(W, H, D, world, C.byref(p), min_planes, min_voxels, frame_ghost,
                                          C.byref(bl), C.byref(br))
    return rc == 0, int(bl.value), int(br.value)


def frame_ghost(W, H, D, world, params=None, min_planes=12, min_voxels=1 << 18):
    """smallest frame ghost (planes) the sharded solve of this geometry needs; 0 = cannot be partitioned"""
    p = make_params(params)
    return int(load().flow3d_sharded_frame_ghost(W, H, D, world, C.byref(p), min_planes, min_voxels))


class ShardedSolver:
    """one rank of a sharded solve (flow3d_sharded_*)"""

    def __init__(self, W, H, D, device, rank, world, uid=None):
        self.L = load()
        self.whd = (W, H, D)
        self.rank, self.world = rank, world
        self.h = _vp()
        idbuf = C.create_string_buffer(uid, ID_BYTES) if uid is not None else None
        check(self.L.flow3d_sharded_create(W, H, D, device, rank, world, idbuf, C.byref(self.h)), "flow3d_sharded_create")

    def destroy(self):
        if self.h:
            self.L.flow3d_sharded_destroy(self.h)
            self.h = _vp()

    def output_planes(self, params):
        a, b = C.c_size_t(), C.c_size_t()
        check(self.L.flow3d_sharded_output_planes(self.h, C.byref(params), C.byref(a), C.byref(b)), "output_planes")
        return int(a.value), int(b.value)

    def set_thresholds(self, min_planes, min_voxels):
        check(self.L.flow3d_sharded_set_thresholds(self.h, min_planes, min_voxels), "set_thresholds")

    def tune(self, params):
        check(self.L.flow3d_sharded_tune(self.h, C.byref(params)), "flow3d_sharded_tune")

    def set_profiling(self, on):
        check(self.L.flow3d_sharded_set_profiling(self.h, 1 if on else 0), "set_profiling")

    def compute(self, raw0, raw1, raw_z0, raw_planes, ld, params, frame_ghost, outs, capacity, stream=None):
        """raw0/raw1/outs: device pointers (ints or c_void_p); returns the delivered plane range (a, b)"""
        a, b = C.c_size_t(), C.c_size_t()
        check(self.L.flow3d_sharded_compute(self.h, _vp(raw0), _vp(raw1), raw_z0, raw_planes, ld, C.byref(params),
                                            frame_ghost, _vp(outs[0]), _vp(outs[1]), _vp(outs[2]), capacity, C.byref(a),
                                            C.byref(b), _vp(stream) if stream else None), "flow3d_sharded_compute")
        return int(a.value), int(b.value)

    def phase_ms(self):
        ms = (C.c_float * 9)()
        check(self.L.flow3d_sharded_phase_ms(self.h, ms), "phase_ms")
        return dict(zip(PHASES, [float(x) for x in ms]))

    def stats(self):
        st = (C.c_double * 8)()
        check(self.L.flow3d_sharded_stats(self.h, st), "stats")
        return dict(zip(STATS, [float(x) for x in st]))


def compute_host(frame_0, frame_1, devices, params=None, persistent=False):
    """Whole-volume solve on several GPUs of this process (one thread + one rank per device inside the
    library).  frame_0/frame_1: numpy (D,H,W).  Returns ([u, v, w], milliseconds)."""
    L = load()
    f0 = np.ascontiguousarray(frame_0, np.float32)
    f1 = np.ascontiguousarray(frame_1, np.float32)
    d, h, w = f0.shape
    outs = [np.zeros_like(f0) for _ in range(3)]
    devs = (C.c_int * len(devices))(*devices)
    ms = C.c_float(0)
    p = make_params(params)
    ptr = lambda a: a.ctypes.data_as(_vp)
    check(L.flow3d_mgpu_compute_host(w, h, d, len(devices), devs, ptr(f0), ptr(f1), C.byref(p), ptr(outs[0]),
                                     ptr(outs[1]), ptr(outs[2]), C.byref(ms), 1 if persistent else 0),
          "flow3d_mgpu_compute_host")
    return outs, float(ms.value)


def release_host_group():
    load().flow3d_mgpu_compute_host(0, 0, 0, 0, None, None, None, None, None, None, None, None, 0)
