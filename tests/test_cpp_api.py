"""The C++ host API that mirrors the reference's classes (include/flow3d/*.h): built into
libflow3d_b200.so and exercised through the two example programs."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

CLI = os.path.join(ROOT, "build", "flow3d_cli")
EXAMPLE = os.path.join(ROOT, "build", "example_main")
META = json.load(open(os.path.join(GOLDEN, "reference_flows.json")))


def test_cli_usage_and_loud_failure_without_device(lib, tmp_path):
    assert os.path.exists(CLI), "build/flow3d_cli missing: run make"
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 2 and "usage:" in r.stdout
    if lib.flow3d_device_count() > 0:
        pytest.skip("a CUDA device is present")
    a = np.zeros((8, 8, 8), np.uint8)
    a.tofile(tmp_path / "a.raw")
    r = subprocess.run([CLI, "--dims", "8", "8", "8", "--frame0", str(tmp_path / "a.raw"), "--frame1",
                        str(tmp_path / "a.raw")], capture_output=True, text=True)
    assert r.returncode == 1 and "no cuda capable devices" in r.stdout.lower()


@pytest.mark.gpu
def test_example_main_reproduces_reference_flow(pair_128, tmp_path):
    """the reference's driver sequence (src/main.cpp:150-185) compiled against our headers"""
    f0, f1 = pair_128
    f0.astype(np.uint8).tofile(tmp_path / "f0.raw")
    f1.astype(np.uint8).tofile(tmp_path / "f1.raw")
    r = subprocess.run([EXAMPLE, str(tmp_path / "f0.raw"), str(tmp_path / "f1.raw"), "128", "128", "128", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Total GPU computation time" in r.stdout
    for c in "uvw":
        got = np.fromfile(tmp_path / ("flow-%s-128-128-128.raw" % c), np.float32)
        assert hashlib.sha256(got.tobytes()).hexdigest() == META["pair128"]["full_sha256"][c]


@pytest.mark.gpu
def test_cli_f32_vtk_and_params(tmp_path):
    rng = np.random.default_rng(3)
    a = (rng.random((12, 20, 28)) * 255).astype(np.float32)
    b = np.roll(a, 1, axis=2)
    a.tofile(tmp_path / "a.raw")
    b.tofile(tmp_path / "b.raw")
    r = subprocess.run([CLI, "--dims", "28", "20", "12", "--frame0", str(tmp_path / "a.raw"), "--frame1",
                        str(tmp_path / "b.raw"), "--f32", "--out", str(tmp_path / "o"), "--vtk", str(tmp_path / "o.vtk"),
                        "--param", "outer_iterations_count=2", "--param", "warp_levels_count=3", "--reps", "2"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("FLOW3D_SOLVE") == 2
    u = np.fromfile(tmp_path / "o_u.raw", np.float32)
    assert u.size == 12 * 20 * 28 and np.isfinite(u).all()
    vtk = open(tmp_path / "o.vtk", "rb").read()
    assert vtk.startswith(b"# vtk DataFile Version 2.0\n3D Vector field computed by GpuFlow3D\nBINARY\n")
    assert b"DIMENSIONS 28 20 12\n" in vtk and len(vtk) > 3 * 4 * u.size
    # wrong size is an error (data3d.cpp:124-131)
    r = subprocess.run([CLI, "--dims", "28", "20", "13", "--frame0", str(tmp_path / "a.raw"), "--frame1",
                        str(tmp_path / "b.raw"), "--f32"], capture_output=True, text=True)
    assert r.returncode == 2 and "wrong dimensions" in r.stdout


@pytest.mark.gpu
def test_cli_diagnostics_and_warped_outputs(gpu, tmp_path):
    """update-norm report, early stopping, registered volume + error map"""
    f0, f1, _ = gpu.ops.synth_pair(40, 36, 32, truth=False)
    f0.tofile(tmp_path / "a.raw")
    f1.tofile(tmp_path / "b.raw")
    base = [CLI, "--dims", "40", "36", "32", "--frame0", str(tmp_path / "a.raw"), "--frame1", str(tmp_path / "b.raw"),
            "--f32", "--param", "warp_levels_count=4", "--param", "warp_scale_factor=0.8", "--param",
            "outer_iterations_count=6"]
    r = subprocess.run(base + ["--out", str(tmp_path / "p"), "--diagnostics", "--warped", str(tmp_path / "w")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("level ")]
    assert len(lines) == 4 and all("outer   6" in l for l in lines)
    warped = np.fromfile(tmp_path / "w_warped.raw", np.float32).reshape(32, 36, 40)
    err = np.fromfile(tmp_path / "w_error.raw", np.float32).reshape(32, 36, 40)
    flow = [np.fromfile(tmp_path / ("p_%s.raw" % c), np.float32).reshape(32, 36, 40) for c in "uvw"]
    # the registered frame is exactly the stage function's warp of frame 1 by the final flow
    want = gpu.ops.warp(f0, f1, flow[0], flow[1], flow[2], (1.0, 1.0, 1.0))
    assert np.array_equal(warped, want)
    assert np.array_equal(err, np.abs(warped - f0))
    # registration does not hurt: the warped frame is no further from frame 0 than frame 1 was (the tiny
    # volume and the 4-level pyramid recover only part of the 6-degree rotation of this pair)
    inner = (slice(4, -4),) * 3
    assert np.abs(warped - f0)[inner].mean() < np.abs(f1 - f0)[inner].mean()
    # a (huge) tolerance stops every level after its first outer iteration
    r = subprocess.run(base + ["--tolerance", "1e9"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("level ")]
    assert len(lines) == 4 and all("outer   1" in l for l in lines)


@pytest.mark.gpu
def test_cli_verbose_prints_the_reference_level_lines(gpu, tmp_path):
    """`silent = false` (optical_flow_e.cpp:270-271): one "Solve level" line per pyramid level"""
    f0, f1, _ = gpu.ops.synth_pair(40, 36, 32, truth=False)
    f0.tofile(tmp_path / "a.raw")
    f1.tofile(tmp_path / "b.raw")
    cmd = [CLI, "--dims", "40", "36", "32", "--frame0", str(tmp_path / "a.raw"), "--frame1", str(tmp_path / "b.raw"),
           "--f32", "--param", "warp_levels_count=4", "--param", "warp_scale_factor=0.8", "--param",
           "outer_iterations_count=2"]
    quiet = subprocess.run(cmd, capture_output=True, text=True)
    loud = subprocess.run(cmd + ["--verbose"], capture_output=True, text=True)
    assert quiet.returncode == 0 and loud.returncode == 0
    assert "Solve level" not in quiet.stdout
    lines = [l for l in loud.stdout.splitlines() if l.startswith("Solve level")]
    assert len(lines) == 4 and lines[-1].startswith("Solve level  0 (  40 x  36 x  32)")


@pytest.mark.gpu
def test_cli_gpus_shards_over_two_devices(gpu, tmp_path):
    """`flow3d_cli --gpus 2` = OpticalFlowE::SetDevices: C++ host -> libflow3d_b200_mgpu.so (threads + NCCL);
    the flow must equal the single-GPU flow byte for byte"""
    if gpu.load().flow3d_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    f0, f1, _ = gpu.ops.synth_pair(72, 40, 96, truth=False)
    f0.tofile(tmp_path / "a.raw")
    f1.tofile(tmp_path / "b.raw")
    base = [CLI, "--dims", "72", "40", "96", "--frame0", str(tmp_path / "a.raw"), "--frame1", str(tmp_path / "b.raw"),
            "--f32", "--param", "warp_levels_count=14", "--param", "outer_iterations_count=3"]
    env = dict(os.environ, FLOW3D_MGPU_MIN_PLANES="8", FLOW3D_MGPU_MIN_VOXELS="1")
    one = subprocess.run(base + ["--out", str(tmp_path / "one")], capture_output=True, text=True)
    two = subprocess.run(base + ["--out", str(tmp_path / "two"), "--gpus", "2", "--reps", "2"], capture_output=True,
                         text=True, env=env)
    assert one.returncode == 0, one.stdout + one.stderr
    assert two.returncode == 0, two.stdout + two.stderr
    assert "Sharding along z over 2 devices." in two.stdout
    for c in "uvw":
        a = open(tmp_path / ("one_%s.raw" % c), "rb").read()
        b = open(tmp_path / ("two_%s.raw" % c), "rb").read()
        assert a == b


@pytest.mark.gpu
def test_cli_pairs_batch_mode_matches_single_pairs(gpu, tmp_path):
    """--pairs PATTERN FIRST LAST: consecutive pairs with the solver kept alive, frame i+1 reused as frame i and
    the next file read during the current solve; every pair's flow must equal the flow of that pair run alone"""
    frames = []
    rng = np.random.default_rng(5)
    base = (rng.random((16, 24, 32)) * 200).astype(np.float32)
    for i in range(4):
        f = np.roll(base, i, axis=2) + np.float32(i)
        f.tofile(tmp_path / ("fr_%04d.raw" % i))
        frames.append(f)
    common = ["--dims", "32", "24", "16", "--f32", "--param", "warp_levels_count=3", "--param", "outer_iterations_count=2"]
    r = subprocess.run([CLI, "--pairs", str(tmp_path / "fr_%04d.raw"), "0", "3", "--out", str(tmp_path / "b")] + common,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("FLOW3D_SOLVE") == 3
    for i in range(3):
        s = subprocess.run([CLI, "--frame0", str(tmp_path / ("fr_%04d.raw" % i)), "--frame1",
                            str(tmp_path / ("fr_%04d.raw" % (i + 1))), "--out", str(tmp_path / ("s%d" % i))] + common,
                           capture_output=True, text=True)
        assert s.returncode == 0, s.stdout + s.stderr
        for c in "uvw":
            assert open(tmp_path / ("b_%d_%s.raw" % (i, c)), "rb").read() == open(tmp_path / ("s%d_%s.raw" % (i, c)), "rb").read()
