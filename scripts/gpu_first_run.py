#!/usr/bin/env python3
"""First GPU session: reference-vs-ours parity on the shipped pairs, golden fixture extraction, and
rough timings of both builds.  Writes everything under gpurun_out/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_flow3d_b200 as pkg  # noqa: E402
from conftest import load_pair_128, load_pair_slab  # noqa: E402
from oracle import ref_runner  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(os.path.join(OUT, "golden"), exist_ok=True)
report = {}


def ours(f0, f1, params=None, reps=1):
    d, h, w = f0.shape
    of = pkg.OpticalFlowE()
    of.silent = True
    assert of.Initialize(pkg.DataSize4(w, h, d))
    p = dict(pkg.DEFAULTS)
    p.update(params or {})
    out = [np.zeros_like(f0) for _ in range(3)]
    times = []
    for _ in range(reps):
        of.ComputeFlow(f0, f1, out[0], out[1], out[2], p)
        times.append(of.last_timing_ms())
    of.Destroy()
    return out, times


def diff_stats(a, b):
    st = {}
    for n, x, y in zip("uvw", a, b):
        d = np.abs(x - y)
        st[n] = dict(max=float(d.max()), mean=float(d.mean()), n_diff=int((x != y).sum()))
    epe = np.sqrt(sum((x - y) ** 2 for x, y in zip(a, b)))
    st["mean_epe_diff"] = float(epe.mean())
    return st


def case(name, f0, f1, u8, guarded_too=False, reps=2):
    print("==== %s %s" % (name, f0.shape), flush=True)
    r = {}
    (ou, ov, ow), ot = ours(f0, f1, reps=reps)
    r["ours_ms"] = ot
    t = time.time()
    ru = None
    try:
        ru, rv, rw, rt, log = ref_runner.run_reference(f0, f1, reps=reps, u8=u8)
        r["ref_seconds"] = rt
        r["ref_wall"] = time.time() - t
        r["diff_vs_ref"] = diff_stats((ou, ov, ow), (ru, rv, rw))
        r["ref_has_nan"] = bool(np.isnan(ru).any() or np.isnan(rv).any() or np.isnan(rw).any())
        np.savez_compressed(os.path.join(OUT, "golden", name + "_ref_full.npz"), u=ru, v=rv, w=rw)
    except Exception as ex:
        r["ref_as_shipped_failed"] = str(ex)[-400:]
    if guarded_too:
        gu, gv, gw, gt, _ = ref_runner.run_reference(f0, f1, reps=reps, u8=u8, guarded=True)
        r["ref_guarded_seconds"] = gt
        r["diff_vs_ref_guarded"] = diff_stats((ou, ov, ow), (gu, gv, gw))
        if ru is not None:
            r["ref_guarded_vs_ref"] = diff_stats((gu, gv, gw), (ru, rv, rw))
        np.savez_compressed(os.path.join(OUT, "golden", name + "_ref_guarded_full.npz"), u=gu, v=gv, w=gw)
    np.savez_compressed(os.path.join(OUT, "golden", name + "_ours_full.npz"), u=ou, v=ov, w=ow)
    print(json.dumps(r, indent=1), flush=True)
    report[name] = r


if __name__ == "__main__":
    which = sys.argv[1:] or ["128", "slab", "synth256"]
    if "128" in which:
        f0, f1 = load_pair_128()
        case("pair128", f0, f1, True)
    if "slab" in which:
        f0, f1 = load_pair_slab()
        case("slab", f0, f1, True, guarded_too=True)
    for n in (64, 256, 512):
        if "synth%d" % n in which:
            f0, f1, truth = pkg.ops.synth_pair(n, n, n)
            case("synth%d" % n, f0, f1, False, reps=1 if n >= 512 else 2)
            ou = np.load(os.path.join(OUT, "golden", "synth%d_ours_full.npz" % n))
            epe = np.sqrt(sum((ou[c] - t) ** 2 for c, t in zip("uvw", truth)))
            report["synth%d" % n]["epe_vs_truth_mean"] = float(epe.mean())
            report["synth%d" % n]["epe_vs_truth_interior"] = float(epe[16:-16, 16:-16, 16:-16].mean())
            if n >= 256:  # too big to bring back
                for f in ("_ref_full.npz", "_ours_full.npz"):
                    os.remove(os.path.join(OUT, "golden", "synth%d%s" % (n, f)))
    json.dump(report, open(os.path.join(OUT, "first_run_report.json"), "w"), indent=1)
    print("REPORT", json.dumps(report))
