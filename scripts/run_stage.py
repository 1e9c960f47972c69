#!/usr/bin/env python3
"""Run one stage kernel repeatedly on an n^3 volume (for ncu captures and quick event timing).

usage: run_stage.py <sweep|phi_ksi|median|resample|warp> [--size 512] [--reps 5]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_flow3d_b200 as pkg  # noqa: E402
from cuda_flow3d_b200._lib import check, f3, sz3  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("stage")
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--dims", type=str, default="")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--alias-f", action="store_true", help="experiment: pass fx/fy also as fz/ft (13 distinct words)")
ap.add_argument("--alias-most", action="store_true", help="experiment: one buffer for all nine static fields (7 distinct words)")
ap.add_argument("--ld-align", type=int, default=0, help="override the row pitch alignment (floats)")
ap.add_argument("--variant", type=int, default=-1, help="sweep: explicit kernel variant (0 register, 1/2 TMA tiles)")
ap.add_argument("--vec", type=int, default=4)
ap.add_argument("--nchunks", type=int, default=0)
ap.add_argument("--ksi", action="store_true", help="sweep: the variant that also computes ksi")
args = ap.parse_args()
L = pkg.load()
pkg.require_device()
import torch  # noqa: E402  (device buffers + events only)

if args.dims:
    W, H, D = [int(x) for x in args.dims.split("x")]
else:
    W = H = D = args.size
ld = int(L.flow3d_aligned_ld(W))
if args.ld_align:
    ld = (W + args.ld_align - 1) // args.ld_align * args.ld_align
n = ld * H * D
dims = sz3((W, H, D))
h = f3((1.0, 1.0, 1.0))
g = torch.Generator(device="cuda").manual_seed(1)


def rnd(scale=1.0, offset=0.0):
    return torch.randn(n, device="cuda", generator=g) * scale + offset


def P(t):
    return C.c_void_p(t.data_ptr())


st = torch.cuda.current_stream()
sp = C.c_void_p(st.cuda_stream)
fx, fy, fz, ft = rnd(5), rnd(5), rnd(5), rnd(10)
u, v, w = rnd(2), rnd(2), rnd(2)
du, dv, dw = rnd(.2), rnd(.2), rnd(.2)
phi, ksi = rnd(0, 1).abs() + 1, rnd(0, 1).abs() + 1
o = [torch.empty(n, device="cuda") for _ in range(4)]


def run():
    if args.stage == "sweep" and args.variant >= 0:
        check(L.flow3d_sweep_shape(P(fx), P(fy), P(fz), P(ft), P(u), P(v), P(w), P(du), P(dv), P(dw), P(phi), P(ksi),
                                   dims, ld, None, h, 7.5, 0.001, P(o[0]), P(o[1]), P(o[2]), P(o[3]) if args.ksi else None,
                                   args.variant, args.vec, args.nchunks, sp), "sweep_shape")
        return 52.0
    if args.stage == "sweep":
        if args.alias_most:
            check(L.flow3d_sweep(P(phi), P(phi), P(phi), P(phi), P(phi), P(phi), P(phi), P(du), P(dv), P(dw), P(phi), P(phi),
                                 dims, ld, h, 7.5, P(o[0]), P(o[1]), P(o[2]), sp), "sweep")
            return 52.0
        if args.alias_f:
            check(L.flow3d_sweep(P(fx), P(fy), P(fx), P(fy), P(u), P(v), P(w), P(du), P(dv), P(dw), P(phi), P(ksi),
                                 dims, ld, h, 7.5, P(o[0]), P(o[1]), P(o[2]), sp), "sweep")
            return 52.0
        check(L.flow3d_sweep(P(fx), P(fy), P(fz), P(ft), P(u), P(v), P(w), P(du), P(dv), P(dw), P(phi), P(ksi),
                             dims, ld, h, 7.5, P(o[0]), P(o[1]), P(o[2]), sp), "sweep")
        return 52.0
    if args.stage == "phi_ksi":
        check(L.flow3d_phi_ksi(P(fx), P(fy), P(fz), P(ft), P(u), P(v), P(w), P(du), P(dv), P(dw), dims, ld, h,
                               0.001, 0.001, P(o[0]), P(o[1]), sp), "phi_ksi")
        return 40.0
    if args.stage == "median":
        check(L.flow3d_median(P(u), P(o[0]), dims, ld, 5, sp), "median")
        return 8.0
    if args.stage == "warp":
        check(L.flow3d_warp_derivatives(P(fx), P(fy), P(u), P(v), P(w), dims, ld, h, P(o[0]), P(o[1]), P(o[2]),
                                        P(o[3]), sp), "warp")
        return 36.0
    if args.stage == "blur":
        check(L.flow3d_gauss_blur(P(u), P(o[0]), P(o[1]), dims, ld, 2.0, sp), "blur")
        return 24.0
    if args.stage == "resample":
        od = sz3((int(W * 0.95), int(H * 0.95), int(D * 0.95)))
        check(L.flow3d_resample(P(u), dims, ld, P(o[0]), od, int(L.flow3d_aligned_ld(int(W * 0.95))), P(o[1]), P(o[2]),
                                sp), "resample")
        return 8.0
    raise SystemExit("unknown stage")


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(args.reps):
    bpv = run()
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
print("%s %dx%dx%d: %.3f ms/launch-group, %.1f GB/s algorithmic (%.0f B/voxel)" %
      (args.stage, W, H, D, ms, bpv * W * H * D / ms / 1e6, bpv))
