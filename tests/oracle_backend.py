"""TEST INFRASTRUCTURE ONLY: the CPU oracle's slab functions behind the backend interface of
cuda_flow3d_b200/dist.py, so that the partitioning / exchange / out-of-core orchestration can be checked on
CPU (gloo, world_size 2-4) against the oracle's whole-volume solve.  Never imported by the package."""
import numpy as np
import torch

from cuda_flow3d_b200.dist import Slab


class OracleBackend:
    """the CPU test oracle's slab functions behind the same interface (gloo tests only)"""
    name = "oracle"

    def __init__(self, oracle):
        self.o = oracle
        self.dev = torch.device("cpu")
        import ctypes as C_
        L = oracle.lib
        f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        lg, sz, fl = C_.c_long, C_.c_size_t, C_.c_float
        L.o_warp_slab.argtypes = [f32p, f32p, lg, f32p, f32p, f32p, sz, sz, sz, lg, lg, lg, fl, fl, fl, f32p]
        L.o_phi_ksi_slab.argtypes = [f32p] * 8 + [sz, sz, lg, lg, lg, lg] + [fl] * 5 + [f32p] * 2
        L.o_sweep_slab.argtypes = [f32p] * 10 + [sz, sz, lg, lg, lg, lg] + [fl] * 4 + [f32p] * 3
        L.o_median_slab.argtypes = [f32p, f32p, sz, sz, lg, lg, lg, lg, sz]
        L.o_median_slab.restype = C_.c_int
        L.o_resample_z_slab.argtypes = [f32p, sz, sz, lg, lg, f32p, lg, lg, lg, lg]

    def ld(self, w):
        return w

    def empty(self, w, h, dl):
        return torch.empty((dl, h, w), dtype=torch.float32)

    def zeros(self, w, h, dl):
        return torch.zeros((dl, h, w), dtype=torch.float32)

    def blur(self, full, sigma):
        return torch.from_numpy(self.o.gauss_blur(full.numpy(), sigma))

    def blur_slab(self, raw, sigma, lo, hi):
        import ctypes as C_
        a = raw.t.numpy()
        taps, r = self.o.gauss_taps(sigma)
        t1 = self.o.conv_axis(a, taps, r, 0)
        t2 = self.o.conv_axis(t1, taps, r, 1)
        out = np.zeros_like(a)
        f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        fn = self.o.lib.o_conv_z_slab
        fn.argtypes = [f32p, f32p, C_.c_size_t, C_.c_size_t, C_.c_long, C_.c_long, C_.c_long, C_.c_long, f32p, C_.c_int]
        d, hh, ww = a.shape
        fn(t2, out, ww, hh, raw.A, raw.dg, lo - raw.A, hi - raw.A, np.ascontiguousarray(taps, np.float32), r)
        return Slab(torch.from_numpy(out), raw.A, raw.dg, raw.w)

    def resample(self, src, src_whd_global, out_whd_global, out_A, out_lo, out_hi, out=None):
        iw, ih, idg = src_whd_global
        ow, oh, odg = out_whd_global
        a = src.t.numpy()
        t1 = self.o.resample_axis(a, ow, 0)
        t2 = self.o.resample_axis(t1, oh, 1)
        self.o.lib.o_resample_z_slab(t2, ow, oh, src.A, idg, out.t.numpy(), out.A, odg, out_lo - out.A, out_hi - out.A)
        return out

    def warp_terms(self, f0, f1, u, v, w, h, lo, hi):
        # the oracle recomputes the derivatives inside phi/ksi and the sweep from (f0, warped f1), which
        # need the warped frame one plane beyond [lo, hi): warp the whole buffer
        out = torch.zeros_like(u.t)
        dl, hh, ww = u.t.shape
        self.o.lib.o_warp_slab(f0.t.numpy(), f1.t.numpy(), f1.A, u.t.numpy(), v.t.numpy(), w.t.numpy(), ww, hh, u.dg,
                               u.A, 0, dl, h[0], h[1], h[2], out.numpy())
        return (f0.t, out)

    def phi_ksi(self, terms, u, v, w, du, dv, dw, h, eps_s, eps_d, phi, ksi, lo, hi):
        dl, hh, ww = u.t.shape
        self.o.lib.o_phi_ksi_slab(terms[0].numpy(), terms[1].numpy(), u.t.numpy(), v.t.numpy(), w.t.numpy(), du.numpy(),
                                  dv.numpy(), dw.numpy(), ww, hh, u.A, u.dg, lo - u.A, hi - u.A, h[0], h[1], h[2], eps_s,
                                  eps_d, phi.numpy(), ksi.numpy())

    def sweep(self, terms, u, v, w, d_in, phi, ksi, h, alpha, d_out, lo, hi):
        dl, hh, ww = u.t.shape
        self.o.lib.o_sweep_slab(terms[0].numpy(), terms[1].numpy(), u.t.numpy(), v.t.numpy(), w.t.numpy(),
                                *[t.numpy() for t in d_in], phi.numpy(), ksi.numpy(), ww, hh, u.A, u.dg, lo - u.A,
                                hi - u.A, h[0], h[1], h[2], alpha, *[t.numpy() for t in d_out])

    def outer_iteration(self, terms, u, v, w, d_cur, d_alt, phi, ksi, h, inner, alpha, eps_s, eps_d, lo1, hi1):
        A, B, d = u.A, u.B, u.dg
        self.phi_ksi(terms, u, v, w, d_cur[0], d_cur[1], d_cur[2], h, eps_s, eps_d, phi, ksi, lo1, hi1)
        for j in range(1, inner + 1):
            lo = lo1 if lo1 == 0 else lo1 + j
            hi = hi1 if hi1 == d else hi1 - j
            self.sweep(terms, u, v, w, d_cur, phi, ksi, h, alpha, d_alt, lo, hi)
            d_cur, d_alt = d_alt, d_cur
        return d_cur, d_alt

    def add3(self, flow, d):
        for c in range(3):
            flow[c].t.add_(d[c])  # one rounded fp32 add per element == add_3d.cu:37-40

    def median(self, src, dst_t, radius, lo, hi):
        dl, hh, ww = src.t.shape
        rc = self.o.lib.o_median_slab(src.t.numpy(), dst_t.numpy(), ww, hh, src.A, src.dg, lo - src.A, hi - src.A, radius)
        if rc != 0:
            raise ValueError("unsupported median radius")

    def absmax(self, s):
        return float(s.t.abs().max())

    def from_numpy_full(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, np.float32).copy())

    def to_numpy(self, t, w):
        return t[:, :, :w].contiguous().numpy()
