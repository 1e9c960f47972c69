"""The oracle is the checker every parity claim leans on, so its own index arithmetic is checked too: the same
source built with AddressSanitizer + UndefinedBehaviorSanitizer runs full pyramid solves on shapes that stress
the borders (584x388x5-like thin slab, a 4^3 volume, a steep 0.5 pyramid, median 1/3/5/7); the sanitizers must
stay silent and the flows must equal the production build's bit for bit (-O1 + sanitizers vs -O2: with
-ffp-contract=off the arithmetic does not depend on the optimiser)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SHAPES = [(28, 24, 20), (70, 40, 5), (9, 7, 33), (4, 4, 4), (17, 5, 6)]  # W, H, D
MEDS = [5, 3, 7, 1, 5]


def test_oracle_under_sanitizers_is_silent_and_bit_identical(tmp_path, oracle):
    so = str(tmp_path / "liboracle_san.so")
    exe = str(tmp_path / "driver")
    flags = ["-O1", "-g", "-std=c++17", "-fopenmp", "-mfma", "-ffp-contract=off", "-fno-fast-math",
             "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"]
    if subprocess.call(["g++"] + flags + ["-fPIC", "-shared", "-o", so, os.path.join(ROOT, "oracle", "flow3d_oracle.cpp")],
                       stderr=subprocess.DEVNULL) != 0:
        pytest.skip("toolchain without sanitizer runtimes")
    subprocess.check_call(["g++"] + flags + [os.path.join(ROOT, "tests", "oracle_sanitized_driver.cpp"), "-o", exe, so,
                                             "-Wl,-rpath," + str(tmp_path)])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", OMP_NUM_THREADS="4")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-2000:]
    for c, ((W, H, D), med) in enumerate(zip(SHAPES, MEDS)):
        a = np.fromfile(str(tmp_path / ("case%d.raw" % c)), np.float32).reshape(5, D, H, W)
        params = dict(warp_levels_count=40, warp_scale_factor=0.5 if c == 2 else 0.95, outer_iterations_count=3,
                      inner_iterations_count=5, median_radius=med, gaussian_sigma=0.0 if c == 3 else 2.0)
        ref = oracle.compute_flow(np.ascontiguousarray(a[0]), np.ascontiguousarray(a[1]), params)
        for k in range(3):
            assert np.array_equal(ref[k].view(np.uint32), a[2 + k].view(np.uint32)), "case %d component %d" % (c, k)
