"""GPU parity of every stage of the hot path against the CPU oracle, through the C ABI.

Bar: BIT-EXACT (the kernels and the oracle place every rounding where the reference's PTX does);
np.array_equal treats -0 == +0, which is the only slack.
"""
import numpy as np
import pytest

from conftest import random_fields, smooth_volume

pytestmark = pytest.mark.gpu

# (d, h, w): cubes, odd sizes, thin slabs (depth 4/5 like config 2), widths around the vector/warp edges
SHAPES = [(16, 16, 16), (9, 13, 21), (5, 37, 130), (4, 20, 47), (18, 18, 18), (12, 33, 257), (7, 9, 96)]
H_CASES = [(1.0, 1.0, 1.0), (1.0491803, 1.0491803, 1.25), (7.111111, 6.4, 3.2)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("sigma", [2.0, 0.8])
def test_gauss_blur(gpu, oracle, shape, sigma):
    (a,) = random_fields(shape, 1, 1, 50.0)
    assert np.array_equal(gpu.ops.gauss_blur(a, sigma), oracle.gauss_blur(a, sigma))


@pytest.mark.parametrize("shape,out", [((16, 16, 16), (15, 15, 15)), ((9, 13, 21), (23, 12, 9)),
                                       ((5, 37, 130), (124, 36, 5)), ((20, 24, 28), (7, 6, 5)),
                                       ((4, 20, 47), (47, 20, 4)), ((6, 10, 122), (128, 11, 6))])
def test_resample(gpu, oracle, shape, out):
    (a,) = random_fields(shape, 2, 1, 5.0)
    assert np.array_equal(gpu.ops.resample(a, out), oracle.resample(a, out))


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("h", H_CASES)
def test_warp_and_derivatives(gpu, oracle, shape, h):
    f0, f1 = smooth_volume(shape, 3), smooth_volume(shape, 4)
    u, v, w = random_fields(shape, 5, 3, 3.0)
    u[0, 0, 0] = np.nan  # NaN target -> falls back to frame 0 (registration_3d.cu:51-53)
    v[-1, -1, -1] = 1e9
    ref = oracle.warp(f0, f1, u, v, w, h)
    got = gpu.ops.warp(f0, f1, u, v, w, h)
    assert np.array_equal(got, ref)
    # derivatives from the warped pair; separate and fused paths must agree bit for bit
    d_sep = gpu.ops.derivatives(f0, got, h)
    d_fused = gpu.ops.warp_derivatives(f0, f1, u, v, w, h)
    for a, b in zip(d_sep, d_fused):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("h", H_CASES)
def test_phi_ksi_and_sweep(gpu, oracle, shape, h):
    f0, f1w = smooth_volume(shape, 6), smooth_volume(shape, 7)
    u, v, w = random_fields(shape, 8, 3, 2.0)
    du, dv, dw = random_fields(shape, 9, 3, 0.2)
    fx, fy, fz, ft = gpu.ops.derivatives(f0, f1w, h)
    phi_o, ksi_o = oracle.phi_ksi(f0, f1w, u, v, w, du, dv, dw, h, 0.001, 0.001)
    phi_g, ksi_g = gpu.ops.phi_ksi(fx, fy, fz, ft, u, v, w, du, dv, dw, h, 0.001, 0.001)
    assert np.array_equal(phi_g, phi_o)
    assert np.array_equal(ksi_g, ksi_o)
    out_o = oracle.sweep(f0, f1w, u, v, w, du, dv, dw, phi_o, ksi_o, h, 7.5)
    out_g = gpu.ops.sweep(fx, fy, fz, ft, u, v, w, du, dv, dw, phi_o, ksi_o, h, 7.5)
    for a, b in zip(out_g, out_o):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", [(16, 16, 16), (5, 37, 130), (9, 13, 21)])
def test_solve_level(gpu, oracle, shape):
    h = (1.25, 1.1, 1.0)
    f0, f1w = smooth_volume(shape, 10), smooth_volume(shape, 11)
    u, v, w = random_fields(shape, 12, 3, 1.0)
    fx, fy, fz, ft = gpu.ops.derivatives(f0, f1w, h)
    ref = oracle.solve_level(f0, f1w, u, v, w, h, 3, 5, 7.5, 0.001, 0.001)
    got = gpu.ops.solve_level(fx, fy, fz, ft, u, v, w, h, 3, 5, 7.5, 0.001, 0.001)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    # odd sweep count exercises the "bring the iterate home" copy
    ref = oracle.solve_level(f0, f1w, u, v, w, h, 1, 3, 7.5, 0.001, 0.001)
    got = gpu.ops.solve_level(fx, fy, fz, ft, u, v, w, h, 1, 3, 7.5, 0.001, 0.001)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("radius", [1, 3, 4, 5, 7])
def test_median(gpu, oracle, shape, radius):
    (a,) = random_fields(shape, 13, 1, 4.0)
    a[a > 3.0] = 3.0  # ties
    assert np.array_equal(gpu.ops.median(a, radius), oracle.median(a, radius))


def test_median_bad_radius(gpu):
    (a,) = random_fields((8, 8, 8), 14, 1)
    with pytest.raises(gpu.Flow3DError):
        gpu.ops.median(a, 9)


def test_add3(gpu):
    f = random_fields((6, 10, 19), 15, 6)
    got = gpu.ops.add3(*f)
    for i in range(3):
        assert np.array_equal(got[i], f[i] + f[i + 3])
