"""Every launch shape / kernel variant of the Jacobi sweep must give the SAME BITS: the register-marching
warp kernel (vec 1/2/4) and the TMA-staged tile kernels (64x8, 32x16 tiles), for any z chunking, with and
without the fused ksi computation, on whole volumes and on z-slabs with ghost planes.  Checked against
the CPU oracle (solve_3d.cu:425-506 restated) and against the whole-volume result for slabs."""
import ctypes as C

import numpy as np
import pytest

from conftest import random_fields, smooth_volume

pytestmark = pytest.mark.gpu

# (d, h, w): tiny, odd, wider than one 64-column tile, taller than one tile, width % 4 != 0, thin slabs
SHAPES = [(16, 16, 16), (9, 13, 21), (5, 37, 130), (4, 20, 47), (12, 33, 257), (7, 9, 96), (21, 70, 66),
          (10, 19, 191), (6, 8, 64), (33, 24, 72)]
VARIANTS = [(0, 4), (0, 2), (0, 1), (1, 4), (2, 4)]  # (variant, vec)


def _setup(gpu, oracle, shape, h, seed=0):
    f0, f1w = smooth_volume(shape, 6 + seed), smooth_volume(shape, 7 + seed)
    u, v, w = random_fields(shape, 8 + seed, 3, 2.0)
    du, dv, dw = random_fields(shape, 9 + seed, 3, 0.2)
    fx, fy, fz, ft = gpu.ops.derivatives(f0, f1w, h)
    phi, ksi = oracle.phi_ksi(f0, f1w, u, v, w, du, dv, dw, h, 0.001, 0.001)
    ref = oracle.sweep(f0, f1w, u, v, w, du, dv, dw, phi, ksi, h, 7.5)
    return (fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi), ref


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("variant,vec", VARIANTS)
def test_sweep_shapes_match_oracle(gpu, oracle, shape, variant, vec):
    h = (1.0491803, 1.0491803, 1.25)
    fields, ref = _setup(gpu, oracle, shape, h)
    ksi = fields[11]
    for nchunks in (0, 2, 5):
        for fused_ksi in (False, True):
            got = gpu.ops.sweep_shape(*fields, h, 7.5, 0.001, variant, vec, nchunks, fused_ksi)
            if got is None:
                pytest.skip("variant %d cannot run %r" % (variant, shape))
            for a, b in zip(got[:3], ref):
                assert np.array_equal(a, b), "variant %d vec %d chunks %d ksi %d" % (variant, vec, nchunks, fused_ksi)
            if fused_ksi:
                assert np.array_equal(got[3], ksi)


@pytest.mark.parametrize("variant", [1, 2])
def test_sweep_tma_sign_of_zero_and_nan(gpu, variant):
    """z-invariant data makes every w-component quantity an exact signed zero (the shipped 584x388x5 pair is
    like that): the TMA kernel must reproduce the register kernel's zero SIGNS (byte equality), and a NaN
    must spread exactly as far."""
    shape = (6, 40, 150)
    h = (1.0, 1.0, 1.25)
    plane0, plane1 = smooth_volume((1,) + shape[1:], 3), smooth_volume((1,) + shape[1:], 4)
    f0 = np.ascontiguousarray(np.broadcast_to(plane0, shape))
    f1 = np.ascontiguousarray(np.broadcast_to(plane1, shape))
    (pu, pv) = random_fields((1,) + shape[1:], 5, 2, 1.0)
    u = np.ascontiguousarray(np.broadcast_to(pu, shape))
    v = np.ascontiguousarray(np.broadcast_to(pv, shape))
    w = np.zeros(shape, np.float32)
    du, dv, dw = np.zeros(shape, np.float32), np.zeros(shape, np.float32), np.zeros(shape, np.float32)
    fx, fy, fz, ft = gpu.ops.derivatives(f0, f1, h)
    phi, ksi = gpu.ops.phi_ksi(fx, fy, fz, ft, u, v, w, du, dv, dw, h, 0.001, 0.001)
    fields = [fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi]
    ref = gpu.ops.sweep_shape(*fields, h, 7.5, 0.001, 0, 4, 0, False)
    got = gpu.ops.sweep_shape(*fields, h, 7.5, 0.001, variant, 4, 3, False)
    for a, b in zip(got[:3], ref[:3]):
        assert a.tobytes() == b.tobytes()
    du2 = du.copy()
    du2[3, 17, 64] = np.nan
    fields[7] = du2
    ref = gpu.ops.sweep_shape(*fields, h, 7.5, 0.001, 0, 4, 0, False)
    got = gpu.ops.sweep_shape(*fields, h, 7.5, 0.001, variant, 4, 2, False)
    for a, b in zip(got[:3], ref[:3]):
        assert np.array_equal(a, b, equal_nan=True) and np.isnan(a).sum() == np.isnan(b).sum()


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("shape", [(24, 13, 21), (30, 20, 130)])
def test_sweep_tma_slabs(gpu, oracle, shape, variant):
    """z-slabs with ghost planes: boundary conditions at the GLOBAL faces only"""
    from cuda_flow3d_b200._lib import ZSlab
    h = (1.05, 1.1, 1.25)
    d = shape[0]
    fields, ref = _setup(gpu, oracle, shape, h, seed=3)
    got = [np.zeros_like(ref[0]) for _ in range(3)]
    m = d // 2
    for (a, b, A, B) in [(0, m, 0, m + 2), (m, d, m - 3, d)]:
        local = [np.ascontiguousarray(x[A:B]) for x in fields]
        out = gpu.ops.sweep_shape(*local, h, 7.5, 0.001, variant, 4, 2, False, slab=ZSlab(A, d, a - A, b - A))
        for c in range(3):
            got[c][a:b] = out[c][a - A:b - A]
    for c in range(3):
        assert np.array_equal(got[c], ref[c])


def test_tuner_is_explicit_and_persisted(gpu, tmp_path, monkeypatch):
    """launches never tune; flow3d_tune_kernels fills the table, which flow3d_tune_query then shows"""
    from cuda_flow3d_b200._lib import check, f3, load, sz3
    L = load()
    dims = (70, 44, 23)
    ld = int(L.flow3d_aligned_ld(dims[0]))
    out = (C.c_int * 3)()
    before = L.flow3d_tune_query(0, sz3(dims), ld, None, out)
    n = ld * dims[1] * dims[2]
    scratch = gpu.DeviceVolume((ld, dims[1], dims[2] * 16), ld=ld)
    L.flow3d_reset_launch_count()
    check(L.flow3d_tune_kernels(sz3(dims), ld, None, f3((1.0, 1.0, 1.0)), scratch.ptr, 16 * n, None), "tune")
    assert L.flow3d_launch_count() == 0, "tuning launches must not be counted as the caller's"
    assert L.flow3d_tune_query(0, sz3(dims), ld, None, out) == 1 or before == 1
    assert out[0] in (1, 2, 4) and 0 <= out[2] <= 2
