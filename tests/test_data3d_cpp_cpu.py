"""Host data formats: the C++ Data3D of this repo against the REFERENCE's own Data3D
(src/data_types/data3d.cpp:95-264).  tests/golden/data3d/ holds what the reference's class, compiled from
/root/reference by scripts/make_data3d_golden.sh, wrote for the operation list tests/data3d_ops.inc: RAW u8
(clamp to [0,255] then truncate; NaN -> 0), RAW f32, the binary VTK vector file, re-reads, Swap, and reads of
files that are too short / too long / missing (refused, volume left empty).  The same list runs here on
include/flow3d/data3d.h; every file must match byte for byte and every outcome must agree.  No GPU needed:
without a device Data3D falls back from page-locked to ordinary host memory."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "data3d")
FILES = ["out_u8.raw", "out_f32.raw", "out_flow.vtk", "reread_u8_as_f32.raw", "reread_f32.raw", "swapped.raw"]


@pytest.fixture(scope="module")
def replay(tmp_path_factory):
    pkg = os.path.join(ROOT, "cuda_flow3d_b200")
    if not os.path.exists(os.path.join(pkg, "libflow3d_b200.so")):
        pytest.fail("libflow3d_b200.so missing: run `make`")
    d = tmp_path_factory.mktemp("data3d")
    exe = str(d / "data3d_check")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests"),
           os.path.join(ROOT, "tests", "data3d_check.cpp"), "-o", exe, "-L" + pkg, "-lflow3d_b200", "-Wl,-rpath," + pkg]
    # address + undefined-behaviour sanitizers when the toolchain has them: the volumes of the failed reads are
    # deleted at the end of the run (the reference's class double-frees there; this one must not)
    if subprocess.call(cmd + ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) != 0:
        subprocess.check_call(cmd)
    out = d / "run"
    out.mkdir()
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0")
    r = subprocess.run([exe, str(out)], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return str(out), r.stdout


@pytest.mark.parametrize("name", FILES)
def test_written_files_equal_the_reference_classes_bytes(replay, name):
    out, _ = replay
    mine = open(os.path.join(out, name), "rb").read()
    ref = open(os.path.join(GOLD, name), "rb").read()
    assert mine == ref, "%s differs from the file the reference's Data3D wrote" % name


def test_outcomes_equal_the_reference_classes(replay):
    out, stdout = replay
    mine = json.load(open(os.path.join(out, "results.json")))
    ref = json.load(open(os.path.join(GOLD, "results.json")))
    assert mine == ref
    # the reference's messages (data3d.cpp:122-131, 50-51)
    assert "wrong dimensions" in stdout and "Cannot open file" in stdout and "Cannot swap two Data3D objects" in stdout


def test_u8_fixture_shows_clamp_truncate_and_nan_rule():
    """what the golden u8 file pins: -0.5 -> 0, 0.49 -> 0, 127.5 -> 127, 254.999 -> 254, 255.5 -> 255, 1e9 -> 255,
    NaN -> 0 (std::max(0.f, NaN) keeps its first argument), +inf -> 255, -inf -> 0, 0.999999 -> 0, 1.0 -> 1"""
    b = open(os.path.join(GOLD, "out_u8.raw"), "rb").read()
    assert len(b) == 7 * 5 * 3
    assert list(b[:12]) == [0, 0, 127, 254, 255, 255, 0, 0, 255, 0, 0, 1]


def test_python_mirror_writes_the_same_files(tmp_path, capsys):
    """cuda_flow3d_b200.api.Data3D (the ctypes-side mirror tests and bench.py use) against the same golden files"""
    import numpy as np
    from cuda_flow3d_b200.api import Data3D
    W, H, D = 7, 5, 3
    a = Data3D()
    assert a.ReadRAWFromFileF32(os.path.join(GOLD, "out_f32.raw"), W, H, D)
    assert a.WriteRAWToFileU8(str(tmp_path / "u8.raw")) and a.WriteRAWToFileF32(str(tmp_path / "f32.raw"))
    assert open(tmp_path / "u8.raw", "rb").read() == open(os.path.join(GOLD, "out_u8.raw"), "rb").read()
    assert open(tmp_path / "f32.raw", "rb").read() == open(os.path.join(GOLD, "out_f32.raw"), "rb").read()
    b = Data3D()
    assert b.ReadRAWFromFileU8(os.path.join(GOLD, "out_u8.raw"), W, H, D)
    assert b.DataPtr().tobytes() == open(os.path.join(GOLD, "reread_u8_as_f32.raw"), "rb").read()
    ref = json.load(open(os.path.join(GOLD, "results.json")))
    c = Data3D()
    assert c.ReadRAWFromFileU8(os.path.join(GOLD, "out_u8.raw"), W, H, D + 1) == bool(ref["read_u8_file_too_short"])
    assert c.ReadRAWFromFileU8(os.path.join(GOLD, "out_u8.raw"), W, H, D - 1) == bool(ref["read_u8_file_too_long"])
    assert c.ReadRAWFromFileF32(os.path.join(GOLD, "out_f32.raw"), W + 1, H, D) == bool(ref["read_f32_file_too_short"])
    assert c.ReadRAWFromFileF32(os.path.join(GOLD, "out_f32.raw"), W, H - 1, D) == bool(ref["read_f32_file_too_long"])
    assert c.ReadRAWFromFileU8(str(tmp_path / "missing.raw"), W, H, D) == bool(ref["read_missing_file"])
    small = Data3D(2, 2, 2)
    small.Swap(a)
    assert small.Width() == 2 and a.Width() == W and "Cannot swap" in capsys.readouterr().out
