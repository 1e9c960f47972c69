"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/flow3d_c.h declares, its host-only arithmetic agrees with the oracle, and compute entry points
fail loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "flow3d_c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(flow3d_[a-z0-9_]+)\s*\(", text))
    names -= {"flow3d_level_callback"}
    return sorted(names)


def test_library_exports_every_declared_symbol(lib):
    import cuda_flow3d_b200._lib as L
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert sorted(L.SIGNATURES) == declared, "python binding and header disagree"


def test_version_and_status_strings(lib):
    assert lib.flow3d_version() >= 100
    assert lib.flow3d_status_string(0) == b"ok"
    assert b"no CPU fallback" in lib.flow3d_status_string(-4)


def test_default_params_match_reference_main(lib):
    import cuda_flow3d_b200 as pkg
    p = pkg.Params()
    lib.flow3d_default_params(C.byref(p))
    got = {k: getattr(p, k) for k, _ in pkg.Params._fields_}
    for k, v in pkg.DEFAULTS.items():  # src/main.cpp:77-85
        assert got[k] == pytest.approx(v)


@pytest.mark.parametrize("dims", [(128, 128, 128), (584, 388, 5), (512, 512, 512), (1024, 1024, 1024),
                                  (450, 180, 450), (37, 5, 91), (4, 4, 4)])
@pytest.mark.parametrize("scale", [0.95, 0.5, 0.8, 1.0])
def test_level_schedule_matches_oracle(lib, oracle, dims, scale):
    import cuda_flow3d_b200 as pkg
    W, H, D = dims
    assert lib.flow3d_max_warp_level(W, H, D, scale) == oracle.max_warp_level(W, H, D, scale)
    a = pkg.level_schedule(W, H, D, scale, 40)
    b = oracle.level_schedule(W, H, D, scale, 40)
    assert len(a) == len(b)
    for (la, da, ha), (lb, db, hb) in zip(a, b):
        assert la == lb and da == db
        assert np.array_equal(np.array(ha, np.float32), np.array(hb, np.float32))


def test_level_counts_of_the_baseline_configs(oracle):
    # 128^3 -> 40 levels, coarsest 18^3; slab -> 10 levels, coarsest 369x245x4
    s = oracle.level_schedule(128, 128, 128, 0.95, 40)
    assert len(s) == 40 and s[0][1] == (18, 18, 18) and s[-1][1] == (128, 128, 128)
    s = oracle.level_schedule(584, 388, 5, 0.95, 40)
    assert len(s) == 10 and s[0][1] == (369, 245, 4) and s[-1][1] == (584, 388, 5)
    s = oracle.level_schedule(512, 512, 512, 0.95, 40)
    assert len(s) == 40 and s[0][1] == (70, 70, 70)


def test_no_device_means_loud_failure_not_cpu_fallback(lib):
    if lib.flow3d_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert lib.flow3d_solver_create(32, 32, 32, 0, C.byref(h)) == -4
    import cuda_flow3d_b200 as pkg
    with pytest.raises(pkg.Flow3DError):
        pkg.require_device()
    of = pkg.OpticalFlowE()
    with pytest.raises(pkg.Flow3DError):
        of.Initialize(pkg.DataSize4(16, 16, 16))


def test_argument_checks_do_not_need_a_device(lib):
    dims = (C.c_size_t * 3)(8, 8, 8)
    h = (C.c_float * 3)(1, 1, 1)
    # null pointers / misaligned pitch are rejected before any CUDA call
    assert lib.flow3d_median(None, None, dims, 8, 5, None) == -1
    assert lib.flow3d_warp(None, None, None, None, None, dims, 8, h, None, None) == -1
    fake = C.c_void_p(0x1000)
    assert lib.flow3d_median(fake, C.c_void_p(0x2000), dims, 6, 5, None) == -1  # ld % 4 != 0
    assert lib.flow3d_level_geometry(8, 8, 8, 0.95, -1, dims, h) == -1


def test_data3d_raw_roundtrip(tmp_path):
    import cuda_flow3d_b200 as pkg
    d = pkg.Data3D(5, 4, 3)
    d.DataPtr()[...] = np.arange(60, dtype=np.float32).reshape(3, 4, 5) * 7 - 30
    f32 = str(tmp_path / "a_f32.raw")
    u8 = str(tmp_path / "a_u8.raw")
    assert d.WriteRAWToFileF32(f32) and d.WriteRAWToFileU8(u8)
    e = pkg.Data3D()
    assert e.ReadRAWFromFileF32(f32, 5, 4, 3)
    assert np.array_equal(e.DataPtr(), d.DataPtr())
    assert e.Data(2, 1, 1) == d.DataPtr()[1, 1, 2]
    assert e.ReadRAWFromFileU8(u8, 5, 4, 3)
    assert np.array_equal(e.DataPtr(), np.clip(d.DataPtr(), 0, 255).astype(np.uint8).astype(np.float32))
    assert not e.ReadRAWFromFileF32(f32, 5, 4, 4)  # size mismatch is an error (data3d.cpp:124-131)
    assert not e.ReadRAWFromFileU8(str(tmp_path / "missing.raw"), 5, 4, 3)


def test_max_warp_level_equals_the_reference_function(lib):
    """flow3d_max_warp_level against the REFERENCE's own OpticalFlowBase::GetMaxWarpLevel
    (src/optical_flow/optical_flow_base.cpp:31-56) on 31 449 (W, H, D, scale) cases: the golden table was printed
    by the reference function itself, compiled from /root/reference (scripts/make_levels_golden.sh).  The count
    decides the level list of every solve, so it has to agree exactly -- including scale >= 1 and dims < 4."""
    import ctypes as C
    import lzma
    import struct
    path = os.path.join(ROOT, "tests", "golden", "max_warp_level.txt.xz")
    lib.flow3d_max_warp_level.restype = C.c_size_t
    lib.flow3d_max_warp_level.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_float]
    n = bad = 0
    first = None
    with lzma.open(path, "rt") as f:
        for line in f:
            w, h, d, bits, want = (int(x) for x in line.split())
            scale = struct.unpack("<f", struct.pack("<I", bits))[0]
            got = int(lib.flow3d_max_warp_level(w, h, d, scale))
            n += 1
            if got != want:
                bad += 1
                first = first or (w, h, d, scale, want, got)
    assert n > 30000
    assert bad == 0, "%d of %d differ; first: %s" % (bad, n, first)


def test_gauss_taps_equal_the_reference_function(lib, oracle):
    """the blur taps of the library (flow3d_gauss_taps) and of the oracle against the REFERENCE's own
    ComputeGaussianKernel (cuda_operation_convolution.cpp:85-108): tests/golden/gauss_taps.txt was printed by
    that function, compiled from /root/reference (scripts/make_taps_golden.sh).  Bit-exact, radius included."""
    import ctypes as C
    import struct
    n = 0
    for line in open(os.path.join(ROOT, "tests", "golden", "gauss_taps.txt")):
        v = [int(x) for x in line.split()]
        sigma = struct.unpack("<f", struct.pack("<I", v[0]))[0]
        radius, want = v[1], v[2:]
        assert len(want) == 2 * radius + 1
        o_taps, o_r = oracle.gauss_taps(sigma)
        assert o_r == radius and [int(x) for x in o_taps.view(np.uint32)] == want, "oracle, sigma %g" % sigma
        buf = (C.c_float * 65)()
        r = C.c_size_t(0)
        rc = lib.flow3d_gauss_taps(sigma, buf, 65, C.byref(r))
        if radius > 32:
            assert rc == -2  # FLOW3D_ERR_UNSUPPORTED, the kernels' limit
        else:
            assert rc == 0 and r.value == radius
            got = np.frombuffer(buf, np.float32, 2 * radius + 1).view(np.uint32)
            assert [int(x) for x in got] == want, "library, sigma %g" % sigma
        n += 1
    assert n == 15
