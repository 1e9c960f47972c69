"""ctypes binding of the CPU test oracle (oracle/flow3d_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference legs.  The product package (cuda_flow3d_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_szp = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")

DEFAULTS = dict(  # src/main.cpp:77-85
    warp_levels_count=40,
    warp_scale_factor=0.95,
    outer_iterations_count=40,
    inner_iterations_count=5,
    equation_alpha=7.5,
    equation_smoothness=0.001,
    equation_data=0.001,
    median_radius=5,
    gaussian_sigma=2.0,
)


class OracleParams(C.Structure):
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


LEVEL_CB = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_float),
                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


def build(variant=""):
    target = "liboracle%s.so" % variant
    subprocess.check_call(["make", "-s", "-C", _HERE, target])
    return os.path.join(_HERE, target)


class Oracle:
    def __init__(self, variant=""):
        path = os.path.join(_HERE, "liboracle%s.so" % variant)
        src = os.path.join(_HERE, "flow3d_oracle.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build(variant)
        L = self.lib = C.CDLL(path)
        L.o_num_threads.restype = C.c_int
        L.o_set_num_threads.argtypes = [C.c_int]
        L.o_max_warp_level.restype = C.c_size_t
        L.o_max_warp_level.argtypes = [C.c_size_t] * 3 + [C.c_float]
        L.o_level_geometry.argtypes = [C.c_size_t] * 3 + [C.c_float, C.c_int, _szp, _f32p]
        L.o_gauss_taps.restype = C.c_int
        L.o_gauss_taps.argtypes = [C.c_float, _f32p]
        L.o_conv_axis.argtypes = [_f32p, _f32p] + [C.c_size_t] * 3 + [_f32p, C.c_int, C.c_int]
        L.o_gauss_blur.argtypes = [_f32p, _f32p, _f32p] + [C.c_size_t] * 3 + [C.c_float]
        L.o_resample_axis.argtypes = [_f32p, _szp, _f32p, C.c_size_t, C.c_int]
        L.o_resample.argtypes = [_f32p, _szp, _f32p, _szp, _f32p, _f32p]
        L.o_warp.argtypes = [_f32p] * 5 + [C.c_size_t] * 3 + [C.c_float] * 3 + [_f32p]
        L.o_phi_ksi.argtypes = [_f32p] * 8 + [C.c_size_t] * 3 + [C.c_float] * 5 + [_f32p] * 2
        L.o_sweep.argtypes = [_f32p] * 10 + [C.c_size_t] * 3 + [C.c_float] * 4 + [_f32p] * 3
        L.o_solve_level.argtypes = ([_f32p] * 9 + [C.c_size_t] * 3 + [C.c_float] * 3 +
                                    [C.c_size_t] * 2 + [C.c_float] * 3)
        L.o_add.argtypes = [_f32p, _f32p, C.c_size_t]
        L.o_median.restype = C.c_int
        L.o_median.argtypes = [_f32p, _f32p] + [C.c_size_t] * 4
        L.o_compute_flow.restype = C.c_int
        L.o_compute_flow.argtypes = ([_f32p, _f32p] + [C.c_size_t] * 3 + [C.POINTER(OracleParams)] +
                                     [_f32p] * 3 + [LEVEL_CB, C.c_void_p])

    # ---- helpers; volumes are numpy float32 arrays of shape (D, H, W) -----------------------
    @staticmethod
    def _dims(a):
        d, h, w = a.shape
        return w, h, d

    def num_threads(self):
        return self.lib.o_num_threads()

    def set_num_threads(self, n):
        self.lib.o_set_num_threads(int(n))

    def max_warp_level(self, W, H, D, scale):
        return int(self.lib.o_max_warp_level(W, H, D, scale))

    def level_geometry(self, W, H, D, scale, level):
        dims = np.zeros(3, np.uint64)
        h = np.zeros(3, np.float32)
        self.lib.o_level_geometry(W, H, D, scale, level, dims, h)
        return tuple(int(x) for x in dims), tuple(np.float32(x) for x in h)

    def level_schedule(self, W, H, D, scale, levels):
        top = min(levels, self.max_warp_level(W, H, D, scale)) - 1
        return [(lv,) + self.level_geometry(W, H, D, scale, lv) for lv in range(top, -1, -1)]

    def gauss_taps(self, sigma):
        t = np.zeros(128, np.float32)
        r = self.lib.o_gauss_taps(sigma, t)
        return t[: 2 * r + 1].copy(), r

    def conv_axis(self, a, taps, radius, axis):
        out = np.empty_like(a)
        w, h, d = self._dims(a)
        self.lib.o_conv_axis(a, out, w, h, d, np.ascontiguousarray(taps, np.float32), radius, axis)
        return out

    def gauss_blur(self, a, sigma):
        out = np.empty_like(a)
        tmp = np.empty_like(a)
        w, h, d = self._dims(a)
        self.lib.o_gauss_blur(a, out, tmp, w, h, d, sigma)
        return out

    def resample_axis(self, a, out_n, axis):
        w, h, d = self._dims(a)
        od = [w, h, d]
        od[axis] = out_n
        out = np.empty((od[2], od[1], od[0]), np.float32)
        self.lib.o_resample_axis(a, np.array([w, h, d], np.uint64), out, out_n, axis)
        return out

    def resample(self, a, out_whd):
        w, h, d = self._dims(a)
        ow, oh, od = out_whd
        n = max(w, ow) * max(h, oh) * max(d, od)
        out = np.empty((od, oh, ow), np.float32)
        ta = np.empty(n, np.float32)
        tb = np.empty(n, np.float32)
        self.lib.o_resample(a, np.array([w, h, d], np.uint64), out,
                            np.array([ow, oh, od], np.uint64), ta, tb)
        return out

    def warp(self, f0, f1, u, v, w_, h):
        out = np.empty_like(f0)
        w, hh, d = self._dims(f0)
        self.lib.o_warp(f0, f1, u, v, w_, w, hh, d, h[0], h[1], h[2], out)
        return out

    def phi_ksi(self, f0, f1w, u, v, w_, du, dv, dw, h, eps_s, eps_d):
        phi = np.empty_like(f0)
        ksi = np.empty_like(f0)
        w, hh, d = self._dims(f0)
        self.lib.o_phi_ksi(f0, f1w, u, v, w_, du, dv, dw, w, hh, d, h[0], h[1], h[2], eps_s, eps_d,
                           phi, ksi)
        return phi, ksi

    def sweep(self, f0, f1w, u, v, w_, du, dv, dw, phi, ksi, h, alpha):
        o = [np.empty_like(f0) for _ in range(3)]
        w, hh, d = self._dims(f0)
        self.lib.o_sweep(f0, f1w, u, v, w_, du, dv, dw, phi, ksi, w, hh, d, h[0], h[1], h[2], alpha,
                         o[0], o[1], o[2])
        return o

    def solve_level(self, f0, f1w, u, v, w_, h, outer, inner, alpha, eps_s, eps_d):
        o = [np.zeros_like(f0) for _ in range(3)]
        scratch = np.empty(5 * f0.size, np.float32)
        w, hh, d = self._dims(f0)
        self.lib.o_solve_level(f0, f1w, u, v, w_, o[0], o[1], o[2], scratch, w, hh, d, h[0], h[1],
                               h[2], outer, inner, alpha, eps_s, eps_d)
        return o

    def median(self, a, radius):
        out = np.empty_like(a)
        w, h, d = self._dims(a)
        rc = self.lib.o_median(a, out, w, h, d, radius)
        if rc != 0:
            raise ValueError("unsupported median radius %d" % radius)
        return out

    def compute_flow(self, f0, f1, params=None, level_cb=None):
        p = dict(DEFAULTS)
        p.update(params or {})
        P = OracleParams(**p)
        w, h, d = self._dims(f0)
        o = [np.empty_like(f0) for _ in range(3)]
        if level_cb is not None:
            def _cb(level, dims, pu, pv, pw, _user):
                dw_, dh_, dd_ = dims[0], dims[1], dims[2]
                n = dw_ * dh_ * dd_
                arrs = [np.ctypeslib.as_array(q, shape=(n,)).reshape(dd_, dh_, dw_).copy()
                        for q in (pu, pv, pw)]
                level_cb(level, (dw_, dh_, dd_), *arrs)
            cb = LEVEL_CB(_cb)
        else:
            cb = C.cast(None, LEVEL_CB)
        rc = self.lib.o_compute_flow(np.ascontiguousarray(f0, np.float32),
                                     np.ascontiguousarray(f1, np.float32), w, h, d, C.byref(P),
                                     o[0], o[1], o[2], cb, None)
        if rc != 0:
            raise RuntimeError("oracle compute_flow failed: %d" % rc)
        return o
