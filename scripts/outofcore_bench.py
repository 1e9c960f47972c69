#!/usr/bin/env python3
"""Time the host-resident (out-of-core) solve against the in-core solve on an n^3 synthetic pair.
usage: outofcore_bench.py [--size 256] [--slabs 4] [--concurrency 2] [--outer 40]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_flow3d_b200 as pkg  # noqa: E402
from cuda_flow3d_b200.outofcore import OutOfCoreFlowSolver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--slabs", type=int, default=4)
ap.add_argument("--concurrency", type=int, default=2)
ap.add_argument("--outer", type=int, default=40)
args = ap.parse_args()
pkg.require_device()
n = args.size
f0, f1, _ = pkg.ops.synth_pair(n, n, n, truth=False)
P = dict(pkg.DEFAULTS, outer_iterations_count=args.outer)
s = pkg.OpticalFlowE()
s.silent = True
assert s.Initialize(pkg.DataSize4(n, n, n))
want = [np.zeros_like(f0) for _ in range(3)]
for _ in range(2):
    s.ComputeFlow(f0, f1, want[0], want[1], want[2], P)
t_in = s.last_timing_ms()[0] / 1e3
s.Destroy()
for cache in (True, False):
    ooc = OutOfCoreFlowSolver(slabs=args.slabs, concurrency=args.concurrency, cache_static=cache)
    t = time.perf_counter()
    got = ooc.compute(f0, f1, P)
    dt = time.perf_counter() - t
    same = all(np.array_equal(a, b) for a, b in zip(got, want))
    print("%d^3 outer=%d slabs=%d concurrency=%d cache_static=%s: out-of-core %.2f s (%.2f Mvoxel/s), in-core %.3f s, "
          "identical=%s, H2D %.1f GB, D2H %.1f GB" % (n, args.outer, args.slabs, args.concurrency, cache, dt,
                                                        n ** 3 / dt / 1e6, t_in, same, ooc.stats["h2d_bytes"] / 1e9,
                                                        ooc.stats["d2h_bytes"] / 1e9))
