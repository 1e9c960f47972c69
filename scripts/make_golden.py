#!/usr/bin/env python3
"""Turn the full-resolution reference flows brought back from a GPU run (gpurun_out/golden/*.npz,
written by scripts/gpu_first_run.py) into the small committed fixtures under tests/golden/."""
import hashlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "golden")
OUT = os.path.join(ROOT, "tests", "golden")


def entry(a, source, sub):
    return {"source": source, "subsample": sub,
            "full_sha256": {c: hashlib.sha256(a[c].tobytes()).hexdigest() for c in "uvw"},
            "stats": {c: {"min": float(a[c].min()), "max": float(a[c].max()), "mean": float(a[c].mean()),
                          "abs_mean": float(np.abs(a[c]).mean())} for c in "uvw"}}


meta = {}
a = np.load(os.path.join(SRC, "pair128_ref_full.npz"))
np.savez_compressed(os.path.join(OUT, "pair128_ref_flow_sub4.npz"), **{c: a[c][::4, ::4, ::4].copy() for c in "uvw"})
meta["pair128"] = entry(a, "reference CUDA build (as shipped), oracle/_ref/flow3d_ref on a B200, default "
                           "parameters, u8 input", "[::4,::4,::4]")
a = np.load(os.path.join(SRC, "slab_ref_guarded_full.npz"))
np.savez_compressed(os.path.join(OUT, "slab_ref_guarded_flow_sub4.npz"), **{c: a[c][:, ::4, ::4].copy() for c in "uvw"})
meta["slab"] = entry(a, "reference CUDA build with the guarded kernels (oracle/_ref/kernels_guarded); the "
                        "as-shipped build dies with CUDA_ERROR_ILLEGAL_ADDRESS on this pair on a B200",
                     "[:, ::4, ::4]")
json.dump(meta, open(os.path.join(OUT, "reference_flows.json"), "w"), indent=1)
