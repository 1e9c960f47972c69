#!/usr/bin/env python3
"""Time our solve of the two shipped pairs (BASELINE configs[0], configs[1]); default parameters."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_flow3d_b200 as pkg  # noqa: E402
from conftest import load_pair_128, load_pair_slab  # noqa: E402

pkg.require_device()
for name, (f0, f1) in (("pair128", load_pair_128()), ("slab584x388x5", load_pair_slab())):
    d, h, w = f0.shape
    of = pkg.OpticalFlowE()
    of.silent = True
    assert of.Initialize(pkg.DataSize4(w, h, d))
    out = [np.zeros_like(f0) for _ in range(3)]
    ms = []
    for _ in range(5):
        of.ComputeFlow(f0, f1, out[0], out[1], out[2], dict(pkg.DEFAULTS))
        ms.append(of.last_timing_ms())
    of.Destroy()
    print("%s: total ms (H2D + levels + D2H) %s ; device-only ms %s" %
          (name, ["%.1f" % t[0] for t in ms], ["%.1f" % t[1] for t in ms]))
