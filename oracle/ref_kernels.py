"""Launch the reference's OWN kernels (the PTX under oracle/_ref/kernels, compiled by
oracle/build_ref.sh from /root/reference/src/kernels/*.cu with the reference's flags) one stage at a
time through the CUDA driver API, with the reference's own launch geometry.

TEST INFRASTRUCTURE ONLY.  This is how the CPU oracle is pinned stage by stage against the
reference itself (the reference has no tests or golden vectors of its own).
"""
import ctypes as C
import os

import numpy as np
from cuda.bindings import driver as drv

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def _ck(res):
    err = res[0]
    if err != drv.CUresult.CUDA_SUCCESS:
        raise RuntimeError("CUDA driver error %s" % err)
    return res[1] if len(res) == 2 else res[1:]


class DataSize4(C.Structure):  # src/data_types/data_structs.h:20-25
    _fields_ = [("width", C.c_size_t), ("height", C.c_size_t), ("depth", C.c_size_t), ("pitch", C.c_size_t)]


class RefKernels:
    """All volumes of one RefKernels instance live in containers of identical geometry
    (cw, ch, cd, pitch), like the reference's 15 equal containers (optical_flow_e.cpp:96-111)."""

    def __init__(self, container_whd, guarded=False):
        _ck(drv.cuInit(0))
        dev = _ck(drv.cuDeviceGet(0))
        self.ctx = _ck(drv.cuDevicePrimaryCtxRetain(dev))
        _ck(drv.cuCtxSetCurrent(self.ctx))
        self.cw, self.ch, self.cd = [int(x) for x in container_whd]
        self.pitch_f = (self.cw + 127) // 128 * 128  # floats; 512-byte rows like cuMemAllocPitch
        self.size = DataSize4(self.cw, self.ch, self.cd, self.pitch_f * 4)
        kdir = os.path.join(REF_DIR, "kernels_guarded" if guarded else "kernels")
        self.mod = {}
        for name in ("add_3d", "median_3d", "convolution_3d", "registration_3d", "resample_3d", "solve_3d"):
            with open(os.path.join(kdir, name + ".ptx"), "rb") as f:
                ptx = f.read() + b"\0"
            m = _ck(drv.cuModuleLoadData(ptx))
            self.mod[name] = m
            ptr, nbytes = _ck(drv.cuModuleGetGlobal(m, b"container_size"))
            assert nbytes == C.sizeof(DataSize4)
            _ck(drv.cuMemcpyHtoD(ptr, bytes(self.size), nbytes))
        self._bufs = []

    # ---- containers ---------------------------------------------------------------------------
    def _nbytes(self):
        return self.pitch_f * 4 * self.ch * self.cd

    def alloc(self):
        p = _ck(drv.cuMemAlloc(self._nbytes()))
        _ck(drv.cuMemsetD8(p, 0, self._nbytes()))
        self._bufs.append(p)
        return p

    def upload(self, a):
        """numpy (d,h,w) -> container (top-left-front corner)"""
        a = np.ascontiguousarray(a, np.float32)
        d, h, w = a.shape
        host = np.zeros((self.cd, self.ch, self.pitch_f), np.float32)
        host[:d, :h, :w] = a
        p = self.alloc()
        _ck(drv.cuMemcpyHtoD(p, host.ctypes.data, host.nbytes))
        return p

    def download(self, p, whd):
        w, h, d = whd
        host = np.empty((self.cd, self.ch, self.pitch_f), np.float32)
        _ck(drv.cuMemcpyDtoH(host.ctypes.data, p, host.nbytes))
        return np.ascontiguousarray(host[:d, :h, :w])

    def free_all(self):
        for p in self._bufs:
            drv.cuMemFree(p)
        self._bufs = []

    def _launch(self, module, kernel, grid, block, smem, values, types):
        f = _ck(drv.cuModuleGetFunction(self.mod[module], kernel.encode()))
        values = [int(v) if t is C.c_void_p else v for v, t in zip(values, types)]
        _ck(drv.cuLaunchKernel(f, grid[0], grid[1], grid[2], block[0], block[1], block[2], smem, 0,
                               (tuple(values), tuple(types)), 0))
        _ck(drv.cuCtxSynchronize())

    @staticmethod
    def _grid(whd, block):
        return tuple((int(n) + b - 1) // b for n, b in zip(whd, block))

    # ---- stages, each with the reference's launch geometry -----------------------------------------
    def warp(self, f0, f1, u, v, w, h):
        """registration_3d; block 16x8x4 (cuda_operation_registration.cpp:105-130)"""
        d_, h_, w_ = f0.shape
        whd = (w_, h_, d_)
        ptrs = [self.upload(x) for x in (f0, f1, u, v, w)]
        out = self.alloc()
        block = (16, 8, 4)
        vals = ptrs + [w_, h_, d_, float(h[0]), float(h[1]), float(h[2]), out]
        types = [C.c_void_p] * 5 + [C.c_size_t] * 3 + [C.c_float] * 3 + [C.c_void_p]
        self._launch("registration_3d", "registration_3d", self._grid(whd, block), block, 0, vals, types)
        return self.download(out, whd)

    def phi_ksi(self, f0, f1, u, v, w, du, dv, dw, h, eps_s, eps_d):
        """compute_phi_ksi_3d; block 16x8x4, 8 shared fields (cuda_operation_solve.cpp:142-152,194-221)"""
        d_, h_, w_ = f0.shape
        whd = (w_, h_, d_)
        ptrs = [self.upload(x) for x in (f0, f1, u, v, w, du, dv, dw)]
        phi, ksi = self.alloc(), self.alloc()
        block = (16, 8, 4)
        smem = 18 * 10 * 6 * 4 * 8
        vals = ptrs + [w_, h_, d_, float(h[0]), float(h[1]), float(h[2]), float(eps_s), float(eps_d), phi, ksi]
        types = [C.c_void_p] * 8 + [C.c_size_t] * 3 + [C.c_float] * 5 + [C.c_void_p] * 2
        self._launch("solve_3d", "compute_phi_ksi_3d", self._grid(whd, block), block, smem, vals, types)
        return self.download(phi, whd), self.download(ksi, whd)

    def sweep(self, f0, f1, u, v, w, du, dv, dw, phi, ksi, h, alpha):
        """solve_3d; block 16x8x4, 10 shared fields (cuda_operation_solve.cpp:154-156,223-252)"""
        d_, h_, w_ = f0.shape
        whd = (w_, h_, d_)
        ptrs = [self.upload(x) for x in (f0, f1, u, v, w, du, dv, dw, phi, ksi)]
        outs = [self.alloc() for _ in range(3)]
        block = (16, 8, 4)
        smem = 18 * 10 * 6 * 4 * 10
        vals = ptrs + [w_, h_, d_, float(h[0]), float(h[1]), float(h[2]), float(alpha)] + outs
        types = [C.c_void_p] * 10 + [C.c_size_t] * 3 + [C.c_float] * 4 + [C.c_void_p] * 3
        self._launch("solve_3d", "solve_3d", self._grid(whd, block), block, smem, vals, types)
        return [self.download(o, whd) for o in outs]

    def resample(self, a, out_whd):
        """resample_x/y/z_3d, X -> Y -> Z; block 16x8x8 (cuda_operation_resample.cpp:95-174)"""
        d_, h_, w_ = a.shape
        ow, oh, od = [int(x) for x in out_whd]
        src = self.upload(a)
        out, tmp = self.alloc(), self.alloc()
        block = (16, 8, 8)
        t = [C.c_void_p, C.c_void_p] + [C.c_size_t] * 4
        self._launch("resample_3d", "resample_x_3d", self._grid((ow, h_, d_), block), block, 0,
                     [src, out, ow, h_, d_, w_], t)
        self._launch("resample_3d", "resample_y_3d", self._grid((ow, oh, d_), block), block, 0,
                     [out, tmp, ow, oh, d_, h_], t)
        self._launch("resample_3d", "resample_z_3d", self._grid((ow, oh, od), block), block, 0,
                     [tmp, out, ow, oh, od, d_], t)
        return self.download(out, (ow, oh, od))

    def gauss_blur(self, a, taps, radius):
        """convolutionRows/Columns/SlicesKernel (cuda_operation_convolution.cpp:159-180, grids :193-302).
        `taps` come from the oracle's restatement of ComputeGaussianKernel (host code)."""
        d_, h_, w_ = a.shape
        whd = (w_, h_, d_)
        m = self.mod["convolution_3d"]
        ptr, nbytes = _ck(drv.cuModuleGetGlobal(m, b"c_Kernel"))
        t = np.ascontiguousarray(taps, np.float32)
        _ck(drv.cuMemcpyHtoD(ptr, t.ctypes.data, t.nbytes))
        src = self.upload(a)
        out, tmp = self.alloc(), self.alloc()
        types = [C.c_void_p, C.c_void_p] + [C.c_int] * 5
        tail = [w_, h_, d_, self.pitch_f, int(radius)]
        self._launch("convolution_3d", "convolutionRowsKernel", ((w_ + 63) // 64, (h_ + 3) // 4, (d_ + 3) // 4),
                     (16, 4, 4), 0, [out, src] + tail, types)
        self._launch("convolution_3d", "convolutionColumnsKernel", ((w_ + 3) // 4, (h_ + 63) // 64, (d_ + 3) // 4),
                     (4, 16, 4), 0, [tmp, out] + tail, types)
        self._launch("convolution_3d", "convolutionSlicesKernel", ((w_ + 3) // 4, (h_ + 3) // 4, (d_ + 63) // 64),
                     (4, 4, 16), 0, [out, tmp] + tail, types)
        return self.download(out, whd)

    def median(self, a, radius):
        """median_3d; block 16x8x4 (cuda_operation_median.cpp:106-145)"""
        d_, h_, w_ = a.shape
        whd = (w_, h_, d_)
        src = self.upload(a)
        out = self.alloc()
        block = (16, 8, 4)
        r2 = radius // 2
        smem = (16 + 2 * r2) * (8 + 2 * r2) * (4 + 2 * r2) * 4
        self._launch("median_3d", "median_3d", self._grid(whd, block), block, smem,
                     [src, w_, h_, d_, int(radius), out], [C.c_void_p] + [C.c_size_t] * 4 + [C.c_void_p])
        return self.download(out, whd)

    def add(self, a, b):
        d_, h_, w_ = a.shape
        whd = (w_, h_, d_)
        pa, pb = self.upload(a), self.upload(b)
        block = (16, 8, 4)
        self._launch("add_3d", "add_3d", self._grid(whd, block), block, 0, [pa, pb, w_, h_, d_],
                     [C.c_void_p, C.c_void_p] + [C.c_size_t] * 3)
        return self.download(pa, whd)
