#!/usr/bin/env python3
"""A/B of the programmatic-dependent-launch chain (flow3d_set_pdl) on the shipped pairs and a synthetic 256^3
pair: alternating off/on solves in ONE process on one solver object, device-only and total ms (CUDA events
inside the library), sha256 of the flows of both modes.  Prints one JSON line."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_flow3d_b200 as pkg  # noqa: E402
from conftest import load_pair_128, load_pair_slab  # noqa: E402

pkg.require_device()
L = pkg.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cases = [("shipped 128^3", load_pair_128()), ("shipped 584x388x5", load_pair_slab())]
if len(sys.argv) > 2:
    n = int(sys.argv[2])
    f0, f1, _ = pkg.ops.synth_pair(n, n, n)
    cases.append(("synthetic %d^3" % n, (f0, f1)))
res = []
for name, (f0, f1) in cases:
    d, h, w = f0.shape
    of = pkg.OpticalFlowE()
    of.silent = True
    assert of.Initialize(pkg.DataSize4(w, h, d))
    out = [np.zeros_like(f0) for _ in range(3)]
    P = dict(pkg.DEFAULTS)
    of.ComputeFlow(f0, f1, out[0], out[1], out[2], P)  # warm-up (first-use quick tuning pass)
    t = {0: [], 1: []}
    sha = {}
    for r in range(reps):
        for mode in (0, 1):
            eff = L.flow3d_set_pdl(mode)
            of.ComputeFlow(f0, f1, out[0], out[1], out[2], P)
            assert of.last_status == 0
            t[mode].append(of.last_timing_ms())
            sha[mode] = hashlib.sha256(b"".join(o.tobytes() for o in out)).hexdigest()
            if mode == 1 and eff != 1:
                sha["refused"] = True
    L.flow3d_set_pdl(-1)
    of.Destroy()
    med = lambda xs, i: float(np.median([x[i] for x in xs]))
    res.append({"case": name, "off_total_ms": med(t[0], 0), "on_total_ms": med(t[1], 0),
                "off_device_ms": med(t[0], 1), "on_device_ms": med(t[1], 1),
                "speedup_device": med(t[0], 1) / med(t[1], 1), "bitwise_equal": sha[0] == sha[1],
                "pdl_refused": bool(sha.get("refused", False))})
print(json.dumps({"pdl_ab": res, "reps": reps}))
