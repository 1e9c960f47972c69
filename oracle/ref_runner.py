"""Run the reference's own CUDA build (oracle/_ref, made by oracle/build_ref.sh) on a GPU box.

TEST/BENCH INFRASTRUCTURE ONLY (tests/, bench.py --impl reference).  Nothing here reads
/root/reference: only the prebuilt binary + PTX under oracle/_ref/, which travel with the repo.
"""
import os
import re
import shutil
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "flow3d_ref")) and os.path.isdir(os.path.join(REF_DIR, "kernels"))


def _stage(guarded):
    """Copy binary + PTX to a short path: the reference builds '<exe dir>/kernels/x.ptx' in a
    256-byte buffer (cuda_operation_solve.cpp:48-52)."""
    dst = "/tmp/f3dref_g" if guarded else "/tmp/f3dref"
    if not os.path.exists(os.path.join(dst, "flow3d_ref")):
        os.makedirs(os.path.join(dst, "kernels"), exist_ok=True)
        shutil.copy2(os.path.join(REF_DIR, "flow3d_ref"), os.path.join(dst, "flow3d_ref"))
        src_k = os.path.join(REF_DIR, "kernels_guarded" if guarded else "kernels")
        for f in os.listdir(src_k):
            shutil.copy2(os.path.join(src_k, f), os.path.join(dst, "kernels", f))
    return os.path.join(dst, "flow3d_ref")


def run_reference_files(p0, p1, dims, reps=1, out_prefix=None, params=None, guarded=False, u8=False,
                        timeout=3600):
    """Run the reference build on two RAW files already on disk (dims = (W, H, D)).  Returns
    ([seconds per rep], stdout); the flows are written to <out_prefix>_{u,v,w}.raw when a prefix is given.
    Nothing of this repository's own library is involved: the process started here is oracle/_ref/flow3d_ref."""
    if not available():
        raise RuntimeError("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    exe = _stage(guarded)
    w, h, d = dims
    cmd = [exe, p0, p1, str(w), str(h), str(d), "u8" if u8 else "f32", out_prefix or "-", str(reps)]
    for k, v in (params or {}).items():
        cmd.append("%s=%s" % (k, v))
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, text=True)
    if res.returncode != 0:
        raise RuntimeError("reference run failed (%d):\n%s" % (res.returncode, res.stdout[-2000:]))
    times = [float(x) for x in re.findall(r"REF_SOLVE rep=\d+ seconds=([0-9.]+)", res.stdout)]
    return times, res.stdout


def run_reference(f0, f1, params=None, reps=1, guarded=False, u8=False, want_output=True, timeout=3600,
                  workdir=None):
    """f0, f1: numpy (D,H,W) volumes.  Returns (u, v, w, [seconds per rep], stdout)."""
    if not available():
        raise RuntimeError("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    exe = _stage(guarded)
    d, h, w = f0.shape
    base = workdir or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
    tmp = tempfile.mkdtemp(prefix="f3dref_", dir=base)
    try:
        p0, p1 = os.path.join(tmp, "f0.raw"), os.path.join(tmp, "f1.raw")
        if u8:
            f0.astype(np.uint8).tofile(p0)
            f1.astype(np.uint8).tofile(p1)
        else:
            np.ascontiguousarray(f0, np.float32).tofile(p0)
            np.ascontiguousarray(f1, np.float32).tofile(p1)
        prefix = os.path.join(tmp, "flow") if want_output else "-"
        cmd = [exe, p0, p1, str(w), str(h), str(d), "u8" if u8 else "f32", prefix, str(reps)]
        for k, v in (params or {}).items():
            cmd.append("%s=%s" % (k, v))
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, text=True)
        out = res.stdout
        if res.returncode != 0:
            raise RuntimeError("reference run failed (%d):\n%s" % (res.returncode, out[-2000:]))
        times = [float(x) for x in re.findall(r"REF_SOLVE rep=\d+ seconds=([0-9.]+)", out)]
        flows = [None] * 3
        if want_output:
            flows = [np.fromfile(prefix + "_%s.raw" % c, np.float32).reshape(d, h, w) for c in "uvw"]
        return flows[0], flows[1], flows[2], times, out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
