"""z-slab (multi-GPU) variants of the stage functions on ONE GPU: a volume is cut into two slabs with
ghost planes, each slab goes through the *_slab C-ABI call, and the stitched result must equal the
whole-volume call bit for bit (boundary conditions apply at the global faces only)."""
import ctypes as C

import numpy as np
import pytest

from conftest import random_fields, smooth_volume

pytestmark = pytest.mark.gpu


def _dv(gpu, a):
    return gpu.DeviceVolume.from_numpy(a)


def _slabs(d, ghost):
    """two slabs: owned [0,m), [m,d) with `ghost` planes of overlap"""
    m = d // 2
    return [(0, m, 0, min(d, m + ghost)), (m, d, max(0, m - ghost), d)]  # (a, b, A, B)


@pytest.mark.parametrize("shape", [(24, 13, 21), (18, 20, 130)])
def test_phi_ksi_sweep_median_slabs(gpu, shape):
    from cuda_flow3d_b200._lib import ZSlab, check, f3, load, sz3
    L = load()
    d, h, w = shape
    hh = (1.05, 1.1, 1.25)
    f0, f1w = smooth_volume(shape, 6), smooth_volume(shape, 7)
    u, v, ww = random_fields(shape, 8, 3, 2.0)
    du, dv, dw = random_fields(shape, 9, 3, 0.2)
    fx, fy, fz, ft = gpu.ops.derivatives(f0, f1w, hh)
    phi, ksi = gpu.ops.phi_ksi(fx, fy, fz, ft, u, v, ww, du, dv, dw, hh, 0.001, 0.001)
    ref_sweep = gpu.ops.sweep(fx, fy, fz, ft, u, v, ww, du, dv, dw, phi, ksi, hh, 7.5)
    ref_med = gpu.ops.median(u, 5)
    got_phi, got_ksi = np.zeros_like(phi), np.zeros_like(ksi)
    got_sweep = [np.zeros_like(phi) for _ in range(3)]
    got_med = np.zeros_like(phi)
    for (a, b, A, B) in _slabs(d, 3):
        sl = lambda x: np.ascontiguousarray(x[A:B])
        dvs = [_dv(gpu, sl(x)) for x in (fx, fy, fz, ft, u, v, ww, du, dv, dw)]
        dims = sz3((w, h, B - A))
        ld = dvs[0].ld
        slab = ZSlab(A, d, a - A, b - A)
        o_phi, o_ksi = gpu.DeviceVolume.zeros((w, h, B - A)), gpu.DeviceVolume.zeros((w, h, B - A))
        check(L.flow3d_phi_ksi_slab(*[t.ptr for t in dvs], dims, ld, C.byref(slab), f3(hh), 0.001, 0.001, o_phi.ptr,
                                    o_ksi.ptr, None), "phi_ksi_slab")
        got_phi[a:b], got_ksi[a:b] = o_phi.numpy()[a - A:b - A], o_ksi.numpy()[a - A:b - A]
        # sweep with the full (reference) phi/ksi slab
        p_phi, p_ksi = _dv(gpu, sl(phi)), _dv(gpu, sl(ksi))
        outs = [gpu.DeviceVolume.zeros((w, h, B - A)) for _ in range(3)]
        check(L.flow3d_sweep_slab(*[t.ptr for t in dvs], p_phi.ptr, p_ksi.ptr, dims, ld, C.byref(slab), f3(hh), 7.5,
                                  *[t.ptr for t in outs], None), "sweep_slab")
        for c in range(3):
            got_sweep[c][a:b] = outs[c].numpy()[a - A:b - A]
        o_med = gpu.DeviceVolume.zeros((w, h, B - A))
        check(L.flow3d_median_slab(dvs[4].ptr, o_med.ptr, dims, ld, C.byref(slab), 5, None), "median_slab")
        got_med[a:b] = o_med.numpy()[a - A:b - A]
    assert np.array_equal(got_phi, phi) and np.array_equal(got_ksi, ksi)
    for c in range(3):
        assert np.array_equal(got_sweep[c], ref_sweep[c])
    assert np.array_equal(got_med, ref_med)


@pytest.mark.parametrize("shape", [(24, 13, 21), (18, 20, 70)])
def test_warp_derivatives_and_resample_slabs(gpu, shape):
    from cuda_flow3d_b200._lib import ZSlab, check, f3, load, sz3
    L = load()
    d, h, w = shape
    hh = (1.0, 1.2, 1.5)
    f0, f1 = smooth_volume(shape, 3), smooth_volume(shape, 4)
    u, v, ww = random_fields(shape, 5, 3, 2.0)
    ref = gpu.ops.warp_derivatives(f0, f1, u, v, ww, hh)
    reach = int(np.ceil(np.abs(ww).max() / hh[2])) + 2
    got = [np.zeros_like(f0) for _ in range(4)]
    for (a, b, A, B) in _slabs(d, 2):
        A1, B1 = max(0, A - reach), min(d, B + reach)
        sl = lambda x: np.ascontiguousarray(x[A:B])
        t = [_dv(gpu, sl(x)) for x in (f0, u, v, ww)]
        tf1 = _dv(gpu, np.ascontiguousarray(f1[A1:B1]))
        outs = [gpu.DeviceVolume.zeros((w, h, B - A)) for _ in range(4)]
        slab = ZSlab(A, d, a - A, b - A)
        check(L.flow3d_warp_derivatives_slab(t[0].ptr, tf1.ptr, A1, B1 - A1, t[1].ptr, t[2].ptr, t[3].ptr,
                                             sz3((w, h, B - A)), t[0].ld, C.byref(slab), f3(hh), *[o.ptr for o in outs],
                                             None), "warp_derivatives_slab")
        for c in range(4):
            got[c][a:b] = outs[c].numpy()[a - A:b - A]
    for c in range(4):
        assert np.array_equal(got[c], ref[c])

    # resample: output slabs computed from just the input planes they need
    from cuda_flow3d_b200.dist import source_range
    out_whd = (w + 3, h - 2, d + 5)
    ref_r = gpu.ops.resample(u, out_whd)
    ow, oh, od = out_whd
    got_r = np.zeros_like(ref_r)
    for (a, b) in [(0, od // 2), (od // 2, od)]:
        s_lo, s_hi = source_range(a, b, d, od)
        src = _dv(gpu, np.ascontiguousarray(u[s_lo:s_hi]))
        out = gpu.DeviceVolume.zeros((ow, oh, b - a))
        ta, tb = gpu.DeviceVolume((ow, h, s_hi - s_lo)), gpu.DeviceVolume((ow, oh, s_hi - s_lo))
        in_slab = ZSlab(s_lo, d, 0, s_hi - s_lo)
        out_slab = ZSlab(a, od, 0, b - a)
        check(L.flow3d_resample_slab(src.ptr, sz3((w, h, s_hi - s_lo)), src.ld, C.byref(in_slab), out.ptr,
                                     sz3((ow, oh, b - a)), out.ld, C.byref(out_slab), ta.ptr, tb.ptr, None),
              "resample_slab")
        got_r[a:b] = out.numpy()
    assert np.array_equal(got_r, ref_r)


def test_absmax(gpu):
    from cuda_flow3d_b200._lib import check, load, sz3
    L = load()
    (a,) = random_fields((7, 9, 21), 3, 1, 5.0)
    v = gpu.DeviceVolume.from_numpy(a)
    out = gpu.DeviceVolume.zeros((4, 1, 1))
    check(L.flow3d_absmax(v.ptr, sz3(v.dims), v.ld, out.ptr, None), "absmax")
    assert out.numpy().ravel()[0] == np.abs(a).max()


@pytest.mark.parametrize("case", [(40, 13, 21, 6, 30), (40, 13, 21, 0, 26), (44, 20, 70, 14, 44)])
def test_outer_iteration_early_late_split(gpu, case):
    """flow3d_outer_iteration_slab_part: early + late = the unsplit iteration bit for bit, and the EARLY part
    must not read the ghost planes at all: it runs with the ghosts of the starting iterate poisoned (NaN),
    the ghosts are restored (the "exchange arrives"), then the late part runs."""
    from cuda_flow3d_b200._lib import ZSlab, check, f3, load, sz3
    L = load()
    d, h, w, a, b = case  # level depth; this rank owns global planes [a, b)
    H = 6
    A, B = max(0, a - H), min(d, b + H)
    hh = (1.05, 1.1, 1.25)
    shape = (d, h, w)
    f0, f1w = smooth_volume(shape, 6), smooth_volume(shape, 7)
    u, v, ww = random_fields(shape, 8, 3, 2.0)
    du, dv, dw = random_fields(shape, 9, 3, 0.2)
    fx, fy, fz, ft = gpu.ops.derivatives(f0, f1w, hh)
    sl = lambda x: np.ascontiguousarray(x[A:B])
    dims = sz3((w, h, B - A))
    lo1 = A if A == 0 else A + 1
    hi1 = B if B == d else B - 1
    slab = ZSlab(A, d, lo1 - A, hi1 - A)

    def run(parts, poison):
        st = [_dv(gpu, sl(x)) for x in (fx, fy, fz, ft, u, v, ww)]
        cur = [sl(x).copy() for x in (du, dv, dw)]
        good = [c.copy() for c in cur]
        if poison:
            for c in cur:
                c[:a - A] = np.nan
                c[b - A:] = np.nan
        dcur = [_dv(gpu, c) for c in cur]
        dalt = [gpu.DeviceVolume.zeros((w, h, B - A)) for _ in range(3)]
        phi, ksi = gpu.DeviceVolume.zeros((w, h, B - A)), gpu.DeviceVolume.zeros((w, h, B - A))
        flag = C.c_int(0)
        for part in parts:
            if part == 2 and poison:  # the exchange arrives: ghosts of the starting iterate become valid
                for t, gd in zip(dcur, good):
                    fixed = t.numpy()
                    fixed[:a - A] = gd[:a - A]
                    fixed[b - A:] = gd[b - A:]
                    check(L.flow3d_upload(fixed.ctypes.data_as(C.c_void_p), t.ptr, dims, t.ld, None), "upload")
                    check(L.flow3d_stream_synchronize(None), "sync")
            check(L.flow3d_outer_iteration_slab_part(*[t.ptr for t in st], *[t.ptr for t in dcur], *[t.ptr for t in dalt],
                                                     phi.ptr, ksi.ptr, dims, st[0].ld, C.byref(slab), f3(hh), 5, 7.5, 0.001,
                                                     0.001, part, a - A, b - A, C.byref(flag), None), "outer_iteration_part")
        res = dalt if flag.value else dcur
        return [t.numpy() for t in res], phi.numpy(), ksi.numpy()

    ref, ref_phi, ref_ksi = run([0], False)
    got, got_phi, got_ksi = run([1, 2], True)
    # valid after 5 sweeps: everything at least H-1 planes away from a non-face buffer end
    v_lo = 0 if A == 0 else 6
    v_hi = (B - A) if B == d else (B - A) - 6
    for c in range(3):
        assert np.array_equal(got[c][v_lo:v_hi], ref[c][v_lo:v_hi])
        assert not np.isnan(got[c][v_lo:v_hi]).any()
    assert np.array_equal(got_phi[lo1 - A:hi1 - A], ref_phi[lo1 - A:hi1 - A])
