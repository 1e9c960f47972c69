"""Python mirror of the reference's host interface for the hot path, over the C ABI.

Names follow the reference (src/optical_flow/optical_flow_e.h:38-66, src/data_types/data3d.h:22-62,
src/data_types/operation_parameters.h:24-33, src/data_types/data_structs.h:20-25) so that the
parity tests read like code written against the reference.  Volumes are numpy float32 arrays of
shape (depth, height, width) -- x fastest, the reference's `(z*H + y)*W + x` (data3d.h:30-32).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import Params, check, f3, load, require_device, sz3

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


def make_params(overrides=None):
    p = dict(DEFAULTS)
    if overrides:
        for k, v in overrides.items():
            if k not in p:
                raise KeyError("unknown solver parameter %r" % k)
            p[k] = v
    return Params(**p)


def level_schedule(W, H, D, scale, levels):
    """[(level, (w,h,d), (hx,hy,hz))] from the coarsest level down to 0
    (optical_flow_base.cpp:31-56, optical_flow_e.cpp:179-180,262-268)."""
    L = load()
    top = min(levels, L.flow3d_max_warp_level(W, H, D, scale)) - 1
    out = []
    for lv in range(top, -1, -1):
        dims = (C.c_size_t * 3)()
        h = (C.c_float * 3)()
        check(L.flow3d_level_geometry(W, H, D, scale, lv, dims, h), "flow3d_level_geometry")
        out.append((lv, tuple(int(x) for x in dims), tuple(np.float32(x) for x in h)))
    return out


class DataSize4:
    """data_structs.h:20-25"""

    def __init__(self, width=0, height=0, depth=0, pitch=0):
        self.width, self.height, self.depth, self.pitch = int(width), int(height), int(depth), int(pitch)


class OperationParameters:
    """Named parameter bag (operation_parameters.h:24-33).  The reference stores `void*` to
    caller-owned variables; here the bag holds the values."""

    def __init__(self):
        self._map = {}

    def PushValuePtr(self, key, value):
        if key in self._map:
            return False
        self._map[key] = value
        return True

    def GetValuePtr(self, key):
        return self._map.get(key)

    def Clear(self):
        self._map.clear()


class Data3D:
    """Host float volume with RAW I/O (data3d.h:22-62, data3d.cpp:95-231)."""

    def __init__(self, width=0, height=0, depth=0):
        self._a = np.zeros((int(depth), int(height), int(width)), np.float32)

    def Width(self):
        return self._a.shape[2]

    def Height(self):
        return self._a.shape[1]

    def Depth(self):
        return self._a.shape[0]

    def DataPtr(self):
        return self._a

    def Data(self, x, y, z):
        return self._a[z, y, x]

    def Swap(self, other):
        if self._a.shape != other._a.shape:  # data3d.cpp:44-52: same-shape volumes only
            print("Error. Cannot swap two Data3D objects (wrong dimensions).")
            return
        self._a, other._a = other._a, self._a

    def ZeroData(self):
        self._a[...] = 0

    def _read(self, filename, width, height, depth, dtype):
        n = int(width) * int(height) * int(depth)
        try:
            size = os.path.getsize(filename)
        except OSError:
            return False
        if size != n * np.dtype(dtype).itemsize:  # data3d.cpp:124-131: exact size or error
            return False
        raw = np.fromfile(filename, dtype=dtype)
        self._a = np.ascontiguousarray(raw.astype(np.float32).reshape(int(depth), int(height), int(width)))
        return True

    def ReadRAWFromFileU8(self, filename, width, height, depth):
        return self._read(filename, width, height, depth, np.uint8)

    def ReadRAWFromFileF32(self, filename, width, height, depth):
        return self._read(filename, width, height, depth, np.float32)

    def WriteRAWToFileU8(self, filename):
        # data3d.cpp:189-190: min(255, max(0, x)) then truncate; std::max(0.f, NaN) keeps the 0
        a = np.where(np.isnan(self._a), np.float32(0), self._a)
        np.clip(a, 0, 255).astype(np.uint8).tofile(filename)
        return True

    def WriteRAWToFileF32(self, filename):
        self._a.astype(np.float32).tofile(filename)
        return True


class DeviceVolume:
    """A pitched fp32 device volume owned through the C ABI's allocator."""

    def __init__(self, dims, ld=None):
        L = load()
        require_device()
        self.dims = tuple(int(x) for x in dims)  # (w, h, d)
        self.ld = int(ld) if ld else int(L.flow3d_aligned_ld(self.dims[0]))
        self.nbytes = self.ld * self.dims[1] * self.dims[2] * 4
        p = C.c_void_p()
        check(L.flow3d_malloc(C.byref(p), self.nbytes), "flow3d_malloc")
        self.ptr = p

    @classmethod
    def from_numpy(cls, a, ld=None):
        a = np.ascontiguousarray(a, np.float32)
        d, h, w = a.shape
        v = cls((w, h, d), ld)
        check(load().flow3d_memset(v.ptr, 0, v.nbytes, None), "flow3d_memset")
        check(load().flow3d_upload(a.ctypes.data_as(C.c_void_p), v.ptr, sz3(v.dims), v.ld, None), "flow3d_upload")
        check(load().flow3d_stream_synchronize(None), "sync")
        return v

    @classmethod
    def zeros(cls, dims, ld=None):
        v = cls(dims, ld)
        check(load().flow3d_memset(v.ptr, 0, v.nbytes, None), "flow3d_memset")
        return v

    def numpy(self):
        w, h, d = self.dims
        out = np.empty((d, h, w), np.float32)
        check(load().flow3d_download(self.ptr, out.ctypes.data_as(C.c_void_p), sz3(self.dims), self.ld, None),
              "flow3d_download")
        check(load().flow3d_stream_synchronize(None), "sync")
        return out

    def free(self):
        if self.ptr:
            load().flow3d_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _Ops:
    """Stage-level calls on numpy volumes (upload -> C ABI -> download); the form the parity tests
    use.  Each mirrors one reference operation's Execute()."""

    @staticmethod
    def _dv(*arrays):
        return [DeviceVolume.from_numpy(a) for a in arrays]

    def gauss_blur(self, a, sigma):
        L = load()
        (x,) = self._dv(a)
        out, tmp = DeviceVolume(x.dims), DeviceVolume(x.dims)
        check(L.flow3d_gauss_blur(x.ptr, out.ptr, tmp.ptr, sz3(x.dims), x.ld, sigma, None), "flow3d_gauss_blur")
        return out.numpy()

    def resample(self, a, out_whd):
        L = load()
        (x,) = self._dv(a)
        ow, oh, od = [int(t) for t in out_whd]
        out = DeviceVolume((ow, oh, od))
        w, h, d = x.dims
        ta = DeviceVolume((ow, h, d))
        tb = DeviceVolume((ow, oh, d))
        check(L.flow3d_resample(x.ptr, sz3(x.dims), x.ld, out.ptr, sz3(out.dims), out.ld, ta.ptr, tb.ptr, None),
              "flow3d_resample")
        return out.numpy()

    def warp(self, f0, f1, u, v, w, h):
        L = load()
        d = self._dv(f0, f1, u, v, w)
        out = DeviceVolume(d[0].dims)
        check(L.flow3d_warp(*[t.ptr for t in d], sz3(d[0].dims), d[0].ld, f3(h), out.ptr, None), "flow3d_warp")
        return out.numpy()

    def derivatives(self, f0, f1w, h):
        L = load()
        d = self._dv(f0, f1w)
        o = [DeviceVolume(d[0].dims) for _ in range(4)]
        check(L.flow3d_derivatives(d[0].ptr, d[1].ptr, sz3(d[0].dims), d[0].ld, f3(h), *[t.ptr for t in o], None),
              "flow3d_derivatives")
        return [t.numpy() for t in o]

    def warp_derivatives(self, f0, f1, u, v, w, h):
        L = load()
        d = self._dv(f0, f1, u, v, w)
        o = [DeviceVolume(d[0].dims) for _ in range(4)]
        check(L.flow3d_warp_derivatives(*[t.ptr for t in d], sz3(d[0].dims), d[0].ld, f3(h),
                                        *[t.ptr for t in o], None), "flow3d_warp_derivatives")
        return [t.numpy() for t in o]

    def phi_ksi(self, fx, fy, fz, ft, u, v, w, du, dv, dw, h, eps_s, eps_d):
        L = load()
        d = self._dv(fx, fy, fz, ft, u, v, w, du, dv, dw)
        phi, ksi = DeviceVolume(d[0].dims), DeviceVolume(d[0].dims)
        check(L.flow3d_phi_ksi(*[t.ptr for t in d], sz3(d[0].dims), d[0].ld, f3(h), eps_s, eps_d, phi.ptr,
                               ksi.ptr, None), "flow3d_phi_ksi")
        return phi.numpy(), ksi.numpy()

    def sweep(self, fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, h, alpha):
        L = load()
        d = self._dv(fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi)
        o = [DeviceVolume(d[0].dims) for _ in range(3)]
        check(L.flow3d_sweep(*[t.ptr for t in d], sz3(d[0].dims), d[0].ld, f3(h), alpha, *[t.ptr for t in o],
                             None), "flow3d_sweep")
        return [t.numpy() for t in o]

    def sweep_shape(self, fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi, h, alpha, eps_d, variant, vec, nchunks,
                    fused_ksi=False, slab=None):
        """one sweep with an explicit launch shape / kernel variant (flow3d_sweep_shape); returns
        [du, dv, dw, ksi_out or None], or None when the variant cannot run this level"""
        L = load()
        d = self._dv(fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi)
        o = [DeviceVolume.zeros(d[0].dims) for _ in range(3)]
        ko = DeviceVolume.zeros(d[0].dims) if fused_ksi else None
        rc = L.flow3d_sweep_shape(*[t.ptr for t in d], sz3(d[0].dims), d[0].ld, C.byref(slab) if slab else None,
                                  f3(h), alpha, eps_d, *[t.ptr for t in o], ko.ptr if ko else None, variant, vec,
                                  nchunks, None)
        if rc == _lib.ERR_UNSUPPORTED:
            return None
        check(rc, "flow3d_sweep_shape")
        return [t.numpy() for t in o] + [ko.numpy() if ko else None]

    def solve_level(self, fx, fy, fz, ft, u, v, w, h, outer, inner, alpha, eps_s, eps_d):
        L = load()
        d = self._dv(fx, fy, fz, ft, u, v, w)
        o = [DeviceVolume.zeros(d[0].dims) for _ in range(3)]
        wv, hh, dd = d[0].dims
        scratch = DeviceVolume((d[0].ld, hh, dd * 5), ld=d[0].ld)
        check(L.flow3d_solve_level(*[t.ptr for t in d], *[t.ptr for t in o], scratch.ptr, sz3(d[0].dims),
                                   d[0].ld, f3(h), outer, inner, alpha, eps_s, eps_d, None), "flow3d_solve_level")
        return [t.numpy() for t in o]

    def add3(self, u, v, w, du, dv, dw):
        L = load()
        d = self._dv(u, v, w, du, dv, dw)
        check(L.flow3d_add3(*[t.ptr for t in d], sz3(d[0].dims), d[0].ld, None), "flow3d_add3")
        return [t.numpy() for t in d[:3]]

    def median(self, a, radius):
        L = load()
        (x,) = self._dv(a)
        out = DeviceVolume(x.dims)
        check(L.flow3d_median(x.ptr, out.ptr, sz3(x.dims), x.ld, int(radius), None), "flow3d_median")
        return out.numpy()

    def update_norm(self, a, b, z_range=None):
        """(sum |a-b|^2, max |a-b|) over three component pairs (flow3d_update_norm)"""
        L = load()
        d = self._dv(*a, *b)
        out = DeviceVolume((4, 1, 1))
        ws = C.c_void_p()
        check(L.flow3d_malloc(C.byref(ws), L.flow3d_update_norm_workspace_bytes()), "malloc")
        sl = None
        if z_range is not None:
            sl = C.byref(_lib.ZSlab(0, d[0].dims[2], int(z_range[0]), int(z_range[1])))
        check(L.flow3d_update_norm(*[t.ptr for t in d], sz3(d[0].dims), d[0].ld, sl, out.ptr, ws, None), "update_norm")
        raw = out.numpy().reshape(-1).view(np.float64)
        L.flow3d_free(ws)
        return float(raw[0]), float(raw[1])

    def synth_pair(self, W, H, D, seed=20240521, truth=True):
        L = load()
        f0, f1 = DeviceVolume((W, H, D)), DeviceVolume((W, H, D))
        t = [DeviceVolume((W, H, D)) for _ in range(3)] if truth else [None] * 3
        check(L.flow3d_synth_pair(W, H, D, 0, D, f0.ld, seed, f0.ptr, f1.ptr,
                                  *[(x.ptr if x else None) for x in t], None), "flow3d_synth_pair")
        return f0.numpy(), f1.numpy(), [x.numpy() if x else None for x in t]


ops = _Ops()


class OpticalFlowE:
    """Mirror of the reference solver class (optical_flow_e.h:38-66): Initialize(DataSize4),
    ComputeFlow(frame_0, frame_1, flow_u, flow_v, flow_w, params), Destroy(), public `silent`."""

    def __init__(self):
        self._solver = None
        self._size = None
        self.silent = False
        self.last_status = 0
        self._cb_keepalive = None

    def GetName(self):
        return "Optical Flow Single GPU"  # optical_flow_e.cpp:30

    def Initialize(self, data_size, device=0):
        L = load()
        require_device()
        if self._solver:
            self.Destroy()
        h = C.c_void_p()
        st = L.flow3d_solver_create(data_size.width, data_size.height, data_size.depth, device, C.byref(h))
        self.last_status = st
        if st != 0:
            if not self.silent:
                print("Initialization failed: %s %s" % (_lib.status_string(st), _lib.last_cuda_error()))
            return False
        self._solver = h
        self._size = (data_size.width, data_size.height, data_size.depth)
        return True

    def set_level_callback(self, fn):
        """fn(level, (w,h,d), u, v, w) with numpy copies of the level's flow; None to clear."""
        L = load()
        if fn is None:
            self._cb_keepalive = _lib.LEVEL_CALLBACK()
            check(L.flow3d_solver_set_level_callback(self._solver, self._cb_keepalive, None), "set_level_callback")
            return

        def _cb(level, dims, ld, pu, pv, pw, _user):
            d = (dims[0], dims[1], dims[2])
            arrs = []
            for p in (pu, pv, pw):
                out = np.empty((d[2], d[1], d[0]), np.float32)
                check(L.flow3d_download(p, out.ctypes.data_as(C.c_void_p), sz3(d), ld, None), "download")
                check(L.flow3d_stream_synchronize(None), "sync")
                arrs.append(out)
            fn(level, d, *arrs)

        self._cb_keepalive = _lib.LEVEL_CALLBACK(_cb)
        check(L.flow3d_solver_set_level_callback(self._solver, self._cb_keepalive, None), "set_level_callback")

    def ComputeFlow(self, frame_0, frame_1, flow_u, flow_v, flow_w, params):
        """Frames/flows are Data3D (or numpy (D,H,W) float32 arrays); params an OperationParameters
        holding the nine keys of Appendix B (missing key -> message and early return, as
        optical_flow_e.cpp:150-158)."""
        L = load()
        if not self._solver:
            print("Error: '%s' was not initialized." % self.GetName())  # optical_flow_base.cpp:60-66
            self.last_status = _lib.ERR_NOT_INITIALIZED
            return
        vals = {}
        for key in DEFAULTS:
            v = params.GetValuePtr(key) if isinstance(params, OperationParameters) else params.get(key)
            if v is None:
                print("Operation: '%s'. Missing parameter '%s'." % (self.GetName(), key))
                self.last_status = _lib.ERR_INVALID_ARG
                return
            vals[key] = v
        P = Params(**vals)

        def arr(x):
            return x.DataPtr() if isinstance(x, Data3D) else x

        a0 = np.ascontiguousarray(arr(frame_0), np.float32)
        a1 = np.ascontiguousarray(arr(frame_1), np.float32)
        outs = [arr(flow_u), arr(flow_v), arr(flow_w)]
        W, H, D = self._size
        for a in [a0, a1] + outs:
            if a.shape != (D, H, W) or a.dtype != np.float32 or not a.flags.c_contiguous:
                self.last_status = _lib.ERR_INVALID_ARG
                raise ValueError("volumes must be C-contiguous float32 arrays of shape (D,H,W)=%r" % ((D, H, W),))
        st = L.flow3d_solver_compute_host(self._solver, a0.ctypes.data_as(C.c_void_p),
                                          a1.ctypes.data_as(C.c_void_p), C.byref(P),
                                          *[o.ctypes.data_as(C.c_void_p) for o in outs])
        self.last_status = st
        if st != 0:
            raise _lib.Flow3DError(st, "flow3d_solver_compute_host")
        if not self.silent:
            ms = (C.c_float * 2)()
            L.flow3d_solver_last_timing(self._solver, ms)
            print("Total GPU computation time: % 4.4fs" % (ms[0] / 1000.0))  # optical_flow_e.cpp:585

    def set_diagnostics(self, enable=True, update_tolerance=0.0):
        """Record the Jacobi update norm per level and outer iteration (flow3d_solver_set_diagnostics).
        update_tolerance > 0 stops a level early once the RMS update drops below it (non-parity mode)."""
        check(load().flow3d_solver_set_diagnostics(self._solver, 1 if enable else 0, float(update_tolerance)),
              "set_diagnostics")

    def diagnostics(self):
        """[(outer iterations run, rms[...], max_abs[...])] per level, coarsest first"""
        L = load()
        nl, nr = C.c_size_t(0), C.c_size_t(0)
        check(L.flow3d_solver_diagnostics(self._solver, C.byref(nl), None, None, None, 0, C.byref(nr)), "diagnostics")
        cap = max(int(nl.value), int(nr.value), 1)
        per = (C.c_size_t * cap)()
        rms = (C.c_double * cap)()
        mx = (C.c_double * cap)()
        check(L.flow3d_solver_diagnostics(self._solver, C.byref(nl), per, rms, mx, cap, C.byref(nr)), "diagnostics")
        out, k = [], 0
        for l in range(int(nl.value)):
            n = int(per[l])
            out.append((n, [rms[k + i] for i in range(n)], [mx[k + i] for i in range(n)]))
            k += n
        return out

    def last_timing_ms(self):
        ms = (C.c_float * 2)()
        check(load().flow3d_solver_last_timing(self._solver, ms), "last_timing")
        return float(ms[0]), float(ms[1])

    def Destroy(self):
        if self._solver:
            load().flow3d_solver_destroy(self._solver)
            self._solver = None

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass
