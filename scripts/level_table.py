#!/usr/bin/env python3
"""Per-level timing table of the solver kernels over the pyramid of an n^3 solve (default parameters):
where the sweep / phi_ksi time of a whole solve goes, level by level.

usage: level_table.py [--size 512] [--reps 10] [--every 1]
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_flow3d_b200 as pkg  # noqa: E402
from cuda_flow3d_b200._lib import check, f3, sz3  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--every", type=int, default=1)
args = ap.parse_args()
L = pkg.load()
pkg.require_device()
import torch  # noqa: E402

N = args.size
sched = pkg.level_schedule(N, N, N, 0.95, 40)
ldN = int(L.flow3d_aligned_ld(N))
nmax = ldN * N * N
g = torch.Generator(device="cuda").manual_seed(1)
bufs = [torch.randn(nmax, device="cuda", generator=g) for _ in range(12)]
bufs[10] = bufs[10].abs() + 1
bufs[11] = bufs[11].abs() + 1
outs = [torch.empty(nmax, device="cuda") for _ in range(3)]
st = torch.cuda.current_stream()
sp = C.c_void_p(st.cuda_stream)


def P(t):
    return C.c_void_p(t.data_ptr())


tot_sw = tot_pk = 0.0
tot_vox = 0.0
print("level      dims      sweep_us  sweep_GB/s  phiksi_us  phiksi_GB/s")
for idx, (lvl, d, h) in enumerate(sched):
    W, H, D = d
    vox = W * H * D
    tot_vox += vox
    if idx % args.every:
        continue
    ld = int(L.flow3d_aligned_ld(W))
    dims = sz3((W, H, D))
    hh = f3(h)
    fx, fy, fz, ft, u, v, w, du, dv, dw, phi, ksi = bufs

    def sweep(k):
        a = (du, dv, dw) if k % 2 == 0 else tuple(outs)
        b = tuple(outs) if k % 2 == 0 else (du, dv, dw)
        check(L.flow3d_sweep(P(fx), P(fy), P(fz), P(ft), P(u), P(v), P(w), P(a[0]), P(a[1]), P(a[2]), P(phi), P(ksi),
                             dims, ld, hh, 7.5, P(b[0]), P(b[1]), P(b[2]), sp), "sweep")

    def phiksi(k):
        check(L.flow3d_phi_ksi(P(fx), P(fy), P(fz), P(ft), P(u), P(v), P(w), P(du), P(dv), P(dw), dims, ld, hh,
                               0.001, 0.001, P(outs[0]), P(outs[1]), sp), "phi_ksi")

    res = []
    for fn in (sweep, phiksi):
        for k in range(2):
            fn(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for k in range(args.reps):
            fn(k)
        e1.record(st)
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / args.reps * 1e3)
    tot_sw += res[0] * 200 * args.every
    tot_pk += res[1] * 40 * args.every
    print("%3d  %4dx%4dx%4d  %9.1f  %9.1f  %9.1f  %9.1f" %
          (lvl, W, H, D, res[0], 52.0 * vox / res[0] / 1e3, res[1], 40.0 * vox / res[1] / 1e3))
print("estimated per-solve totals: sweep %.1f ms, phi_ksi %.1f ms (x200 / x40 launches per level)" %
      (tot_sw / 1e3, tot_pk / 1e3))
