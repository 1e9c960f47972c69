"""ctypes binding of libflow3d_b200.so (the C ABI declared in include/flow3d_c.h).

There is NO CPU fallback: if the shared library is missing, or no CUDA device is visible when a
compute entry point is called, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflow3d_b200.so")

OK = 0
ERR_INVALID_ARG = -1
ERR_UNSUPPORTED = -2
ERR_CUDA = -3
ERR_NO_DEVICE = -4
ERR_OUT_OF_MEMORY = -5
ERR_NOT_INITIALIZED = -6


class Flow3DError(RuntimeError):
    def __init__(self, status, where=""):
        self.status = status
        msg = "%s: status %d (%s)" % (where, status, status_string(status))
        if status == ERR_CUDA or status == ERR_OUT_OF_MEMORY:
            msg += " -- " + last_cuda_error()
        super().__init__(msg)


class Params(C.Structure):
    """flow3d_params: the nine named parameters of the reference (src/main.cpp:77-85)."""
    _fields_ = [
        ("warp_levels_count", C.c_size_t),
        ("warp_scale_factor", C.c_float),
        ("outer_iterations_count", C.c_size_t),
        ("inner_iterations_count", C.c_size_t),
        ("equation_alpha", C.c_float),
        ("equation_smoothness", C.c_float),
        ("equation_data", C.c_float),
        ("median_radius", C.c_size_t),
        ("gaussian_sigma", C.c_float),
    ]


class ZSlab(C.Structure):
    """flow3d_zslab: z-slab view of a level sharded along z (include/flow3d_c.h)."""
    _fields_ = [("z0_global", C.c_size_t), ("depth_global", C.c_size_t), ("z_begin", C.c_size_t),
                ("z_end", C.c_size_t)]


LEVEL_CALLBACK = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_size_t), C.c_size_t, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_void_p)

_sz3 = C.c_size_t * 3
_f3 = C.c_float * 3
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/flow3d_c.h declares
SIGNATURES = {
    "flow3d_version": (C.c_int, []),
    "flow3d_status_string": (C.c_char_p, [C.c_int]),
    "flow3d_last_cuda_error": (C.c_char_p, []),
    "flow3d_device_count": (C.c_int, []),
    "flow3d_default_params": (None, [C.POINTER(Params)]),
    "flow3d_launch_count": (C.c_uint64, []),
    "flow3d_reset_launch_count": (None, []),
    "flow3d_max_warp_level": (C.c_size_t, [C.c_size_t, C.c_size_t, C.c_size_t, C.c_float]),
    "flow3d_level_geometry": (C.c_int, [C.c_size_t, C.c_size_t, C.c_size_t, C.c_float, C.c_int, _sz3, _f3]),
    "flow3d_aligned_ld": (C.c_size_t, [C.c_size_t]),
    "flow3d_set_device": (C.c_int, [C.c_int]),
    "flow3d_malloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "flow3d_free": (C.c_int, [_vp]),
    "flow3d_memset": (C.c_int, [_vp, C.c_int, C.c_size_t, _vp]),
    "flow3d_upload": (C.c_int, [_vp, _vp, _sz3, C.c_size_t, _vp]),
    "flow3d_download": (C.c_int, [_vp, _vp, _sz3, C.c_size_t, _vp]),
    "flow3d_stream_synchronize": (C.c_int, [_vp]),
    "flow3d_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "flow3d_host_free": (C.c_int, [_vp]),
    "flow3d_device_name": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t]),
    "flow3d_gauss_blur": (C.c_int, [_vp, _vp, _vp, _sz3, C.c_size_t, C.c_float, _vp]),
    "flow3d_resample": (C.c_int, [_vp, _sz3, C.c_size_t, _vp, _sz3, C.c_size_t, _vp, _vp, _vp]),
    "flow3d_warp": (C.c_int, [_vp] * 5 + [_sz3, C.c_size_t, _f3, _vp, _vp]),
    "flow3d_derivatives": (C.c_int, [_vp] * 2 + [_sz3, C.c_size_t, _f3] + [_vp] * 5),
    "flow3d_warp_derivatives": (C.c_int, [_vp] * 5 + [_sz3, C.c_size_t, _f3] + [_vp] * 5),
    "flow3d_phi_ksi": (C.c_int, [_vp] * 10 + [_sz3, C.c_size_t, _f3, C.c_float, C.c_float] + [_vp] * 3),
    "flow3d_sweep": (C.c_int, [_vp] * 12 + [_sz3, C.c_size_t, _f3, C.c_float] + [_vp] * 4),
    "flow3d_solve_level": (C.c_int, [_vp] * 11 + [_sz3, C.c_size_t, _f3, C.c_size_t, C.c_size_t,
                                                  C.c_float, C.c_float, C.c_float, _vp]),
    "flow3d_add3": (C.c_int, [_vp] * 6 + [_sz3, C.c_size_t, _vp]),
    "flow3d_median": (C.c_int, [_vp, _vp, _sz3, C.c_size_t, C.c_size_t, _vp]),
    "flow3d_gauss_blur_slab": (C.c_int, [_vp, _vp, _vp, _sz3, C.c_size_t, C.POINTER(ZSlab), C.c_float, _vp]),
    "flow3d_sweep_slab": (C.c_int, [_vp] * 12 + [_sz3, C.c_size_t, C.POINTER(ZSlab), _f3, C.c_float] + [_vp] * 4),
    "flow3d_sweep_shape": (C.c_int, [_vp] * 12 + [_sz3, C.c_size_t, C.POINTER(ZSlab), _f3, C.c_float, C.c_float] +
                           [_vp] * 4 + [C.c_int] * 3 + [_vp]),
    "flow3d_phi_ksi_slab": (C.c_int, [_vp] * 10 + [_sz3, C.c_size_t, C.POINTER(ZSlab), _f3, C.c_float, C.c_float] + [_vp] * 3),
    "flow3d_warp_derivatives_slab": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, _vp, _vp, _vp, _sz3, C.c_size_t,
                                               C.POINTER(ZSlab), _f3] + [_vp] * 5),
    "flow3d_outer_iteration_slab": (C.c_int, [_vp] * 15 + [_sz3, C.c_size_t, C.POINTER(ZSlab), _f3, C.c_size_t,
                                                C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int), _vp]),
    "flow3d_outer_iteration_slab_part": (C.c_int, [_vp] * 15 + [_sz3, C.c_size_t, C.POINTER(ZSlab), _f3, C.c_size_t,
                                                     C.c_float, C.c_float, C.c_float, C.c_int, C.c_size_t, C.c_size_t,
                                                     C.POINTER(C.c_int), _vp]),
    "flow3d_median_slab": (C.c_int, [_vp, _vp, _sz3, C.c_size_t, C.POINTER(ZSlab), C.c_size_t, _vp]),
    "flow3d_resample_slab": (C.c_int, [_vp, _sz3, C.c_size_t, C.POINTER(ZSlab), _vp, _sz3, C.c_size_t, C.POINTER(ZSlab),
                                       _vp, _vp, _vp]),
    "flow3d_absmax": (C.c_int, [_vp, _sz3, C.c_size_t, _vp, _vp]),
    "flow3d_solver_workspace_bytes": (C.c_size_t, [C.c_size_t] * 3),
    "flow3d_solver_create": (C.c_int, [C.c_size_t] * 3 + [C.c_int, C.POINTER(_vp)]),
    "flow3d_solver_destroy": (C.c_int, [_vp]),
    "flow3d_solver_compute_host": (C.c_int, [_vp, _vp, _vp, C.POINTER(Params), _vp, _vp, _vp]),
    "flow3d_solver_compute_device": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.POINTER(Params), _vp, _vp,
                                               _vp, _vp]),
    "flow3d_solver_tune": (C.c_int, [_vp, C.POINTER(Params)]),
    "flow3d_set_pdl": (C.c_int, [C.c_int]),
    "flow3d_gauss_taps": (C.c_int, [C.c_float, C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_size_t)]),
    "flow3d_tune_kernels": (C.c_int, [_sz3, C.c_size_t, C.POINTER(ZSlab), _f3, _vp, C.c_size_t, _vp]),
    "flow3d_tune_query": (C.c_int, [C.c_int, _sz3, C.c_size_t, C.POINTER(ZSlab), C.c_int * 3]),
    "flow3d_solver_last_timing": (C.c_int, [_vp, C.c_float * 2]),
    "flow3d_solver_set_profiling": (C.c_int, [_vp, C.c_int]),
    "flow3d_solver_set_verbose": (C.c_int, [_vp, C.c_int]),
    "flow3d_solver_stage_times": (C.c_int, [_vp, C.c_float * 8, C.c_double * 8, C.c_uint64 * 8]),
    "flow3d_solver_set_level_callback": (C.c_int, [_vp, LEVEL_CALLBACK, _vp]),
    "flow3d_update_norm_workspace_bytes": (C.c_size_t, []),
    "flow3d_update_norm": (C.c_int, [_vp] * 6 + [_sz3, C.c_size_t, C.POINTER(ZSlab), _vp, _vp, _vp]),
    "flow3d_solver_set_diagnostics": (C.c_int, [_vp, C.c_int, C.c_float]),
    "flow3d_solver_diagnostics": (C.c_int, [_vp, C.POINTER(C.c_size_t), _vp, _vp, _vp, C.c_size_t,
                                            C.POINTER(C.c_size_t)]),
    "flow3d_selftest_fast_div": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int, C.c_uint64 * 3]),
    "flow3d_synth_pair": (C.c_int, [C.c_size_t] * 6 + [C.c_uint64] + [_vp] * 6),
}

_lib = None


def load():
    """Load the C-ABI library (loud failure if it was not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `make` (or __graft_entry__.build()). "
                "cuda_flow3d_b200 has no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def status_string(status):
    return load().flow3d_status_string(status).decode()


def last_cuda_error():
    return load().flow3d_last_cuda_error().decode()


def check(status, where=""):
    if status != OK:
        raise Flow3DError(status, where)


def require_device():
    n = load().flow3d_device_count()
    if n <= 0:
        raise Flow3DError(ERR_NO_DEVICE, "cuda_flow3d_b200 needs a CUDA device")
    return n


def sz3(dims):
    return _sz3(*[int(x) for x in dims])


def f3(h):
    return _f3(*[float(x) for x in h])
