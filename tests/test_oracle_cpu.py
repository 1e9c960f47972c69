"""The CPU oracle against first principles, and against the fixtures produced by the reference's own
CUDA build (tests/golden/README.md)."""
import numpy as np
import pytest

from conftest import random_fields, smooth_volume


def test_gauss_taps(oracle):
    taps, r = oracle.gauss_taps(2.0)
    assert r == 6 and len(taps) == 13  # (size_t)(3*sigma), cuda_operation_convolution.cpp:87
    assert np.array_equal(taps, taps[::-1])
    assert abs(float(taps.sum()) - 1.0) < 1e-6
    assert taps.argmax() == 6


def test_blur_is_zero_padded(oracle):
    a = np.ones((20, 20, 20), np.float32)
    b = oracle.gauss_blur(a, 2.0)
    assert abs(b[10, 10, 10] - 1.0) < 1e-5
    assert b[0, 0, 0] < 0.25  # zero padding darkens the corners (convolution_3d.cu:113,122,130)
    assert np.allclose(b, b[::-1, ::-1, ::-1], atol=1e-6)  # symmetric up to accumulation order


def test_resample_identity_and_mean(oracle):
    (a,) = random_fields((8, 9, 10), 1, 1)
    assert np.array_equal(oracle.resample(a, (10, 9, 8)), a)  # delta == 1: single cell, weight delta
    c = np.full((12, 12, 12), 3.25, np.float32)
    assert np.allclose(oracle.resample(c, (7, 9, 11)), 3.25, atol=2e-6)
    assert np.allclose(oracle.resample(c, (17, 13, 12)), 3.25, atol=2e-6)
    # exact 2:1 box average along x
    d = oracle.resample_axis(a, 5, 0)
    assert np.allclose(d, 0.5 * (a[:, :, 0::2] + a[:, :, 1::2]), atol=1e-6)


def test_warp_zero_flow_and_shift(oracle):
    f0, f1 = smooth_volume((10, 12, 14), 2), smooth_volume((10, 12, 14), 3)
    z = np.zeros_like(f0)
    assert np.array_equal(oracle.warp(f0, f1, z, z, z, (1, 1, 1)), f1)
    u = np.full_like(f0, 2.0)  # integer shift in x with hx = 2 -> one voxel
    out = oracle.warp(f0, f1, u, z, z, (2.0, 1, 1))
    assert np.array_equal(out[:, :, :-1], f1[:, :, 1:])
    assert np.array_equal(out[:, :, -1], f0[:, :, -1])  # target leaves the volume -> frame 0


def test_median_is_a_selection(oracle):
    (a,) = random_fields((7, 8, 9), 4, 1)
    m = oracle.median(a, 5)
    # brute force with mirrored padding
    p = np.pad(a, 2, mode="reflect")
    for (z, y, x) in [(0, 0, 0), (3, 4, 5), (6, 7, 8), (0, 7, 4)]:
        win = np.sort(p[z:z + 5, y:y + 5, x:x + 5].ravel())
        assert m[z, y, x] == win[62]
    assert np.array_equal(oracle.median(a, 1), a)
    assert np.array_equal(oracle.median(a, 4), oracle.median(a, 3))  # even -> radius-1
    with pytest.raises(ValueError):
        oracle.median(a, 9)


def test_sweep_fixed_point_of_constant_fields(oracle):
    # identical frames, zero flow: du stays exactly zero
    f = smooth_volume((8, 9, 10), 5)
    z = np.zeros_like(f)
    phi, ksi = oracle.phi_ksi(f, f, z, z, z, z, z, z, (1, 1, 1), 0.001, 0.001)
    assert np.allclose(phi, 500.0) and np.allclose(ksi, 500.0)  # 1/(2*eps)
    out = oracle.sweep(f, f, z, z, z, z, z, z, phi, ksi, (1, 1, 1), 7.5)
    for o in out:
        assert np.all(o == 0)


def test_oracle_recovers_a_translation(oracle):
    shape = (24, 24, 24)
    d, h, w = shape
    zz, yy, xx = np.meshgrid(np.arange(d), np.arange(h), np.arange(w), indexing="ij")

    def tex(x, y, z):
        return (120 + 40 * np.sin(0.35 * x + 0.3) * np.cos(0.3 * y) + 30 * np.sin(0.28 * z + 0.2 * x)).astype(np.float32)

    f0 = tex(xx, yy, zz)
    f1 = tex(xx - 0.8, yy + 0.5, zz - 0.3)  # f1(x + flow) = f0(x) with flow = (0.8, -0.5, 0.3)
    u, v, w_ = oracle.compute_flow(f0, f1, dict(outer_iterations_count=10, gaussian_sigma=0.5))
    inner = (slice(6, -6),) * 3
    assert abs(u[inner].mean() - 0.8) < 0.15
    assert abs(v[inner].mean() + 0.5) < 0.15
    assert abs(w_[inner].mean() - 0.3) < 0.15


def test_oracle_level_count_equals_the_reference_function(oracle):
    """the oracle's level count against the table printed by the reference's own GetMaxWarpLevel
    (tests/golden/max_warp_level.txt.xz, scripts/make_levels_golden.sh): pins the oracle's schedule to the
    reference's code, not to a reading of it"""
    import lzma
    import os
    import struct
    from conftest import ROOT
    bad = n = 0
    with lzma.open(os.path.join(ROOT, "tests", "golden", "max_warp_level.txt.xz"), "rt") as f:
        for i, line in enumerate(f):
            if i % 3:
                continue
            w, h, d, bits, want = (int(x) for x in line.split())
            scale = struct.unpack("<f", struct.pack("<I", bits))[0]
            n += 1
            bad += int(oracle.max_warp_level(w, h, d, scale) != want)
    assert n > 10000 and bad == 0


def test_warp_agrees_with_the_references_own_cpu_registration(oracle):
    """Cross-check named in SURVEY.md 8c: the reference carries a CPU version of the backward registration
    (cuda_operation_register_p.cpp:96-139).  tests/golden/warp_cpu/ holds its output for a small case with random
    flows plus the special voxels (integer shifts, exact borders, one ulp past the border, NaN, +-inf, -0, 1e9),
    produced by that code compiled from /root/reference (scripts/make_warp_golden.sh).
    * the out-of-volume / NaN fallback (out = frame_0) must hit exactly the same voxels;
    * integer displacements (zero fractions) must agree bit for bit;
    * elsewhere the CPU loop rounds every product and sum separately while the GPU kernel the oracle mirrors
      contracts them into FMAs (registration_3d.cu:66-79 as compiled), so the tolerance is 8 ulp of the
      frames' range (255): 1.3e-4; observed max 6.1e-5, 79 % of the voxels bit-equal."""
    import json
    import os
    from conftest import ROOT
    g = os.path.join(ROOT, "tests", "golden", "warp_cpu")
    c = json.load(open(os.path.join(g, "case.json")))
    W, H, D = c["W"], c["H"], c["D"]

    def ld(name):
        return np.fromfile(os.path.join(g, name + ".raw"), np.float32).reshape(D, H, W)
    f0, f1, u, v, w, ref = (ld(n) for n in ("f0", "f1", "u", "v", "w", "warped_ref_cpu"))
    got = oracle.warp(f0, f1, u, v, w, tuple(c["h"]))
    assert not np.isnan(got).any() and not np.isnan(ref).any()
    assert np.array_equal(got == f0, ref == f0)                      # same fallback voxels
    assert int((got == f0).sum()) == 277
    assert np.array_equal(got[0, 0], ref[0, 0]) and np.array_equal(got[1, 1], ref[1, 1])  # integer shifts
    assert np.abs(got - ref).max() <= 1.3e-4
    assert (got.view(np.uint32) == ref.view(np.uint32)).mean() > 0.7
