"""Pins the CPU oracle, stage by stage, against the reference's OWN kernels: the PTX that
oracle/build_ref.sh compiled from /root/reference/src/kernels/*.cu is loaded through the driver API and
launched with the reference's block/grid geometry on the same seeded inputs.  Bar: bit-exact.
Skipped when oracle/_ref is absent (it is built where /root/reference exists and travels to the box)."""
import os

import numpy as np
import pytest

from conftest import ROOT, random_fields, smooth_volume

pytestmark = pytest.mark.gpu

_HAVE_REF = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "kernels"))
needs_ref = pytest.mark.skipif(not _HAVE_REF, reason="oracle/_ref not built")

# dims are multiples of 4 where the as-shipped blur kernels are involved (they write out of bounds
# otherwise); the other kernels take ragged sizes.
SHAPES = [(16, 16, 16), (9, 13, 21), (8, 24, 40), (5, 37, 70)]
H_CASES = [(1.0, 1.0, 1.0), (1.0491803, 1.0491803, 1.25), (7.111111, 6.4, 3.2)]


@pytest.fixture(scope="module")
def ref(gpu):
    from oracle.ref_kernels import RefKernels
    r = RefKernels((136, 48, 24))
    yield r
    r.free_all()


def _report(name, got, want):
    if not np.array_equal(got, want):
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        rel = d / np.maximum(np.abs(want), 1e-30)
        pytest.fail("%s: %d of %d values differ, max abs %g, max rel %g" %
                    (name, int((got != want).sum()), got.size, d.max(), rel.max()))


@needs_ref
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("h", H_CASES)
def test_ref_warp(ref, oracle, shape, h):
    f0, f1 = smooth_volume(shape, 3), smooth_volume(shape, 4)
    u, v, w = random_fields(shape, 5, 3, 3.0)
    u[0, 0, 0] = np.nan
    _report("warp", oracle.warp(f0, f1, u, v, w, h), ref.warp(f0, f1, u, v, w, h))
    ref.free_all()


@needs_ref
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("h", H_CASES)
def test_ref_phi_ksi(ref, oracle, shape, h):
    f0, f1w = smooth_volume(shape, 6), smooth_volume(shape, 7)
    u, v, w = random_fields(shape, 8, 3, 2.0)
    du, dv, dw = random_fields(shape, 9, 3, 0.2)
    phi_r, ksi_r = ref.phi_ksi(f0, f1w, u, v, w, du, dv, dw, h, 0.001, 0.001)
    phi_o, ksi_o = oracle.phi_ksi(f0, f1w, u, v, w, du, dv, dw, h, 0.001, 0.001)
    _report("phi", phi_o, phi_r)
    _report("ksi", ksi_o, ksi_r)
    ref.free_all()


@needs_ref
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("h", H_CASES)
def test_ref_sweep(ref, oracle, shape, h):
    f0, f1w = smooth_volume(shape, 6), smooth_volume(shape, 7)
    u, v, w = random_fields(shape, 8, 3, 2.0)
    du, dv, dw = random_fields(shape, 9, 3, 0.2)
    phi, ksi = oracle.phi_ksi(f0, f1w, u, v, w, du, dv, dw, h, 0.001, 0.001)
    out_r = ref.sweep(f0, f1w, u, v, w, du, dv, dw, phi, ksi, h, 7.5)
    out_o = oracle.sweep(f0, f1w, u, v, w, du, dv, dw, phi, ksi, h, 7.5)
    for n, a, b in zip(("du", "dv", "dw"), out_o, out_r):
        _report(n, a, b)
    ref.free_all()


@needs_ref
@pytest.mark.parametrize("shape,out", [((16, 16, 16), (15, 15, 15)), ((9, 13, 21), (23, 12, 9)),
                                       ((5, 37, 130), (124, 36, 5)), ((20, 24, 28), (7, 6, 5)),
                                       ((6, 10, 122), (128, 11, 6))])
def test_ref_resample(ref, oracle, shape, out):
    (a,) = random_fields(shape, 2, 1, 5.0)
    _report("resample", oracle.resample(a, out), ref.resample(a, out))
    ref.free_all()


@needs_ref
@pytest.mark.parametrize("shape", [(16, 16, 16), (8, 24, 40), (12, 20, 36)])
@pytest.mark.parametrize("sigma", [2.0, 0.8])
def test_ref_gauss_blur(gpu, oracle, shape, sigma):
    # the blur kernels stride z by imageH (convolution_3d.cu:100), i.e. they assume the container IS
    # the image (true in the reference: the blur only runs at full resolution)
    from oracle.ref_kernels import RefKernels
    d, h, w = shape
    r = RefKernels((w, h, d))
    (a,) = random_fields(shape, 1, 1, 50.0)
    taps, radius = oracle.gauss_taps(sigma)
    try:
        _report("blur", oracle.gauss_blur(a, sigma), r.gauss_blur(a, taps, radius))
    finally:
        r.free_all()


@needs_ref
@pytest.mark.parametrize("shape", [(16, 16, 16), (8, 13, 21), (12, 24, 40)])
@pytest.mark.parametrize("radius", [3, 5, 7])
def test_ref_median(ref, oracle, shape, radius):
    (a,) = random_fields(shape, 13, 1, 4.0)
    a[a > 3.0] = 3.0
    _report("median", oracle.median(a, radius), ref.median(a, radius))
    ref.free_all()


@needs_ref
def test_ref_add(ref):
    a, b = random_fields((9, 13, 21), 15, 2)
    _report("add", a + b, ref.add(a, b))
    ref.free_all()
