"""z-sharded multi-GPU flow solve: one process per GPU, `torch.distributed` for the plumbing.

Decomposition (the reference's own precedent is the z-slab + halo scheme of its
out-of-core path, src/cuda_operations/partial_data/cuda_operation_solve_p.cpp:358-417):

* every level with enough planes is split along z: rank g owns planes [a, b) = [g*d/G, (g+1)*d/G) and
  keeps a buffer [A, B) = [a-H, b+H) (clipped at the global faces) with H = inner_iterations + 1 ghost
  planes per side;
* the solver is point-Jacobi, so with an H-deep ghost zone the phi/ksi update plus all
  `inner` sweeps of one outer iteration run WITHOUT communication on shrinking ranges (phi on
  [A+1,B-1), sweep j on [A+1+j, B-1-j)); one neighbour exchange of H planes of (du,dv,dw) per outer
  iteration restores the ghosts.  The result is bit-identical to the single-GPU solve;
* the blurred full-resolution frames are replicated on every rank (each rank gets both input frames),
  so per-level frames are resampled locally for exactly the planes a rank needs -- including the
  data-dependent z reach of the warp (max|w|/hz planes, from an on-device max reduction);
* the flow is exchanged only between z neighbours: H ghost planes after the prolongation, and the
  per-outer-iteration exchange above; the 5^3 median and the next prolongation read ghosts that are
  already valid.  No collective sits on the data path;
* coarse levels (too few planes to shard) are computed redundantly by every rank (replicas): no
  communication, and every rank holds the full flow when the first sharded level starts.

The compute backend is the C ABI (`CabiBackend`, CUDA tensors).  The backend is duck-typed so that the CPU
tests (tests/oracle_backend.py, gloo, world_size 2-4) can plug the test oracle's slab functions into the
same orchestration; nothing in this package knows the oracle.

Since round 2 the PRODUCT multi-GPU path is the C++ sharded solver (csrc/sharded_solver.cu, bound by
mgpu.py: NCCL send/recv issued from C++ on the solve's streams, ghost exchange overlapped with compute).
This module remains as the orchestration the out-of-core solver (outofcore.py) runs its virtual ranks
through and as the executable specification the gloo tests check the partition arithmetic against.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import ZSlab, check, f3, load, sz3
from .api import DEFAULTS


def own_range(d, rank, world):
    return (rank * d) // world, ((rank + 1) * d) // world


def f32(x):
    return np.float32(x)


def source_range(o_lo, o_hi, a, b):
    """input planes [lo, hi) read by output planes [o_lo, o_hi) of an axis resampled from a to b samples
    (same float arithmetic as the kernels: resample_3d.cu:41-48)"""
    delta = f32(a) / f32(b)
    lo = int(np.floor(f32(o_lo) * delta))
    hi = int(min(f32(a), np.ceil(f32(o_hi) * delta)))
    return max(lo, 0), min(hi, a)


class Slab:
    """a z-slab of one field of one level: tensor (dl, h, ld) holding global planes [A, A+dl)"""

    def __init__(self, t, A, dg, w):
        self.t, self.A, self.dg, self.w = t, int(A), int(dg), int(w)

    @property
    def dl(self):
        return self.t.shape[0]

    @property
    def B(self):
        return self.A + self.t.shape[0]

    def planes(self, g_lo, g_hi):
        return self.t[g_lo - self.A:g_hi - self.A]


# =====================================================================================================
class CabiBackend:
    """stage functions of libflow3d_b200.so on CUDA tensors (current torch stream)"""
    name = "cabi"

    def __init__(self, device):
        self.L = load()
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        check(self.L.flow3d_set_device(device), "set_device")

    def _sp(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def ld(self, w):
        return int(self.L.flow3d_aligned_ld(w))

    def empty(self, w, h, dl):
        return torch.empty((dl, h, self.ld(w)), dtype=torch.float32, device=self.dev)

    def zeros(self, w, h, dl):
        return torch.zeros((dl, h, self.ld(w)), dtype=torch.float32, device=self.dev)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def _slab(self, s, lo, hi):
        return ZSlab(s.A, s.dg, lo - s.A, hi - s.A)

    def blur(self, full, sigma):
        d, h, ld = full.shape
        out, tmp = torch.empty_like(full), torch.empty_like(full)
        check(self.L.flow3d_gauss_blur(self._p(full), self._p(out), self._p(tmp), sz3((self.w_full, h, d)), ld, sigma,
                                       self._sp()), "blur")
        return out

    def blur_slab(self, raw, sigma, lo, hi):
        """raw: Slab of the full-resolution frame; returns a Slab (same planes) whose planes [lo, hi) hold the
        blurred frame (zero padding at the global faces only)"""
        d, h, ld = raw.t.shape
        out, tmp = torch.empty_like(raw.t), torch.empty_like(raw.t)
        sl = ZSlab(raw.A, raw.dg, lo - raw.A, hi - raw.A)
        check(self.L.flow3d_gauss_blur_slab(self._p(raw.t), self._p(out), self._p(tmp), sz3((raw.w, h, d)), ld,
                                            C.byref(sl), sigma, self._sp()), "gauss_blur_slab")
        return Slab(out, raw.A, raw.dg, raw.w)

    def resample(self, src, src_whd_global, out_whd_global, out_A, out_lo, out_hi, out=None):
        """src: Slab of the input level; returns/fills a Slab of the output level with plane 0 = out_A,
        computing global output planes [out_lo, out_hi)"""
        iw, ih, idg = src_whd_global
        ow, oh, odg = out_whd_global
        if out is None:
            raise ValueError("out slab required")
        tmp_a = torch.empty((src.dl, ih, self.ld(ow)), dtype=torch.float32, device=self.dev)
        tmp_b = torch.empty((src.dl, oh, self.ld(ow)), dtype=torch.float32, device=self.dev)
        in_slab = ZSlab(src.A, idg, 0, src.dl)
        out_slab = ZSlab(out.A, odg, out_lo - out.A, out_hi - out.A)
        check(self.L.flow3d_resample_slab(self._p(src.t), sz3((iw, ih, src.dl)), src.t.shape[2], C.byref(in_slab),
                                          self._p(out.t), sz3((ow, oh, out.dl)), out.t.shape[2], C.byref(out_slab),
                                          self._p(tmp_a), self._p(tmp_b), self._sp()), "resample_slab")
        return out

    def warp_terms(self, f0, f1, u, v, w, h, lo, hi):
        """image terms on global planes [lo, hi): (fx, fy, fz, ft) slabs (same geometry as u)"""
        terms = [torch.empty_like(u.t) for _ in range(4)]
        dims = sz3((u.w, u.t.shape[1], u.dl))
        sl = self._slab(u, lo, hi)
        check(self.L.flow3d_warp_derivatives_slab(self._p(f0.t), self._p(f1.t), f1.A, f1.dl, self._p(u.t), self._p(v.t),
                                                  self._p(w.t), dims, u.t.shape[2], C.byref(sl), f3(h),
                                                  *[self._p(t) for t in terms], self._sp()), "warp_derivatives_slab")
        return terms

    def phi_ksi(self, terms, u, v, w, du, dv, dw, h, eps_s, eps_d, phi, ksi, lo, hi):
        dims = sz3((u.w, u.t.shape[1], u.dl))
        sl = self._slab(u, lo, hi)
        check(self.L.flow3d_phi_ksi_slab(*[self._p(t) for t in terms], self._p(u.t), self._p(v.t), self._p(w.t),
                                         self._p(du), self._p(dv), self._p(dw), dims, u.t.shape[2], C.byref(sl), f3(h),
                                         eps_s, eps_d, self._p(phi), self._p(ksi), self._sp()), "phi_ksi_slab")

    def sweep(self, terms, u, v, w, d_in, phi, ksi, h, alpha, d_out, lo, hi):
        dims = sz3((u.w, u.t.shape[1], u.dl))
        sl = self._slab(u, lo, hi)
        check(self.L.flow3d_sweep_slab(*[self._p(t) for t in terms], self._p(u.t), self._p(v.t), self._p(w.t),
                                       *[self._p(t) for t in d_in], self._p(phi), self._p(ksi), dims, u.t.shape[2],
                                       C.byref(sl), f3(h), alpha, *[self._p(t) for t in d_out], self._sp()),
              "sweep_slab")

    def outer_iteration(self, terms, u, v, w, d_cur, d_alt, phi, ksi, h, inner, alpha, eps_s, eps_d, lo1, hi1):
        """phi/ksi on [lo1,hi1) + `inner` sweeps on shrinking ranges in one C call; returns (cur, alt)"""
        dims = sz3((u.w, u.t.shape[1], u.dl))
        sl = self._slab(u, lo1, hi1)
        flag = C.c_int(0)
        check(self.L.flow3d_outer_iteration_slab(*[self._p(t) for t in terms], self._p(u.t), self._p(v.t), self._p(w.t),
                                                 *[self._p(t) for t in d_cur], *[self._p(t) for t in d_alt],
                                                 self._p(phi), self._p(ksi), dims, u.t.shape[2], C.byref(sl), f3(h),
                                                 inner, alpha, eps_s, eps_d, C.byref(flag), self._sp()),
              "outer_iteration_slab")
        return (d_alt, d_cur) if flag.value else (d_cur, d_alt)

    def add3(self, flow, d):
        u = flow[0]
        check(self.L.flow3d_add3(self._p(flow[0].t), self._p(flow[1].t), self._p(flow[2].t), self._p(d[0]),
                                 self._p(d[1]), self._p(d[2]), sz3((u.w, u.t.shape[1], u.dl)), u.t.shape[2], self._sp()),
              "add3")

    def median(self, src, dst_t, radius, lo, hi):
        dims = sz3((src.w, src.t.shape[1], src.dl))
        sl = self._slab(src, lo, hi)
        check(self.L.flow3d_median_slab(self._p(src.t), self._p(dst_t), dims, src.t.shape[2], C.byref(sl), radius,
                                        self._sp()), "median_slab")

    def absmax(self, s):
        out = torch.zeros(1, dtype=torch.float32, device=self.dev)
        check(self.L.flow3d_absmax(self._p(s.t), sz3((s.w, s.t.shape[1], s.dl)), s.t.shape[2], self._p(out),
                                   self._sp()), "absmax")
        return float(out.item())

    def from_numpy_full(self, a):
        """tight numpy (D,H,W) -> padded device tensor (D,H,ld)"""
        d, h, w = a.shape
        t = self.zeros(w, h, d)
        t[:, :, :w] = torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.w_full = w
        return t

    def to_numpy(self, t, w):
        return t[:, :, :w].contiguous().cpu().numpy()


# =====================================================================================================
class TorchComm:
    """the three communication patterns of the sharded solve on torch.distributed (NCCL on GPUs, gloo in
    the CPU tests): neighbour send/recv, one scalar max, one all-gather of equal-sized pieces"""

    def __init__(self, rank=None, world=None):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world

    def sendrecv(self, ops):
        """ops: list of ("send" | "recv", tensor, peer), matched in order per peer"""
        p2p = [dist.P2POp(dist.isend if kind == "send" else dist.irecv, t, peer) for kind, t, peer in ops]
        for r in dist.batch_isend_irecv(p2p):
            r.wait()

    def start_sendrecv(self, ops):
        """post the operations and return at once (NCCL: they run on the communicator's stream after
        the work already queued on the current stream); finish_sendrecv() orders later work after them"""
        p2p = [dist.P2POp(dist.isend if kind == "send" else dist.irecv, t, peer) for kind, t, peer in ops]
        return dist.batch_isend_irecv(p2p) if p2p else []

    def finish_sendrecv(self, handle):
        for r in handle:
            r.wait()

    def allreduce_max(self, value, device):
        t = torch.tensor([value], dtype=torch.float32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_gather(self, piece):
        pieces = [torch.empty_like(piece) for _ in range(self.world)]
        dist.all_gather(pieces, piece)
        return pieces


class LocalComm:
    """the same patterns between `world` threads of ONE process (the out-of-core solver's virtual ranks):
    per-pair FIFO mailboxes for send/recv, a barrier for the two collectives"""

    class Shared:
        def __init__(self, world):
            import queue
            import threading
            self.world = world
            self.box = {(s, d): queue.Queue() for s in range(world) for d in range(world)}
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world
            self.failed = threading.Event()  # set by a rank that died: the others stop waiting for it

    def __init__(self, shared, rank):
        self.sh, self.rank, self.world = shared, rank, shared.world

    def sendrecv(self, ops):
        for kind, t, peer in ops:  # post every send first: nobody blocks on a full mailbox
            if kind == "send":
                self.sh.box[(self.rank, peer)].put(t.clone())
        import queue
        for kind, t, peer in ops:
            if kind == "recv":
                while True:
                    try:
                        t.copy_(self.sh.box[(peer, self.rank)].get(timeout=0.5))
                        break
                    except queue.Empty:
                        if self.sh.failed.is_set():
                            raise RuntimeError("a peer slab failed") from None

    def start_sendrecv(self, ops):
        for kind, t, peer in ops:
            if kind == "send":
                self.sh.box[(self.rank, peer)].put(t.clone())
        return [op for op in ops if op[0] == "recv"]

    def finish_sendrecv(self, handle):
        self.sendrecv(handle)

    def _exchange_slot(self, value):
        self.sh.slots[self.rank] = value
        self.sh.barrier.wait(timeout=600)
        out = list(self.sh.slots)
        self.sh.barrier.wait(timeout=600)
        return out

    def allreduce_max(self, value, device):
        return float(max(self._exchange_slot(float(value))))

    def all_gather(self, piece):
        return [p.clone() for p in self._exchange_slot(piece)]


# =====================================================================================================
class ReplicatedFrames:
    """both blurred full-resolution frames live on every rank"""
    global_reach = False

    def __init__(self, solver, F0, F1, whd):
        self.s, self.whd = solver, whd
        W, Hh, D = whd
        self.full = [Slab(F0, 0, D, W), Slab(F1, 0, D, W)]

    def level_frame(self, which, level, dims, lo, hi):
        if level == 0:
            return Slab(self.full[which].planes(lo, hi), lo, dims[2], dims[0])
        return self.s._frame(self.full[which], self.whd, dims, lo, hi)


class ShardedFrames:
    """each rank holds only a z-slab [Va, Vb) of the blurred full-resolution frames (its own planes plus
    `ghost` planes per side).  A level frame is resampled locally when every rank's request fits its
    slab; otherwise (coarse levels, whose source intervals are long) every rank resamples the planes whose
    source interval starts in its own range and the level frame is assembled with one all-gather."""
    global_reach = True

    def __init__(self, solver, raw0, raw1, raw_z0, whd, sigma, ghost):
        self.s, self.whd, self.ghost = solver, whd, ghost
        be = solver.be
        W, Hh, D = whd
        fa, fb = own_range(D, solver.rank, solver.world)
        self.Va, self.Vb = max(0, fa - ghost), min(D, fb + ghost)
        self.slabs = []
        for raw in (raw0, raw1):
            rs = Slab(raw, raw_z0, D, W)
            if sigma > 0:
                r = int(3 * sigma)
                assert rs.A <= max(0, self.Va - r) and rs.B >= min(D, self.Vb + r), "raw slab lacks blur halo planes"
                bl = be.blur_slab(rs, sigma, self.Va, self.Vb)
            else:
                bl = rs
            self.slabs.append(Slab(bl.planes(self.Va, self.Vb), self.Va, D, W))
        self._gathered = {}

    @staticmethod
    def input_planes(D, rank, world, sigma, ghost):
        """planes [lo, hi) of the raw frames a rank has to be given"""
        fa, fb = own_range(D, rank, world)
        r = int(3 * sigma) if sigma > 0 else 0
        return max(0, fa - ghost - r), min(D, fb + ghost + r)

    def _fits_everywhere(self, dims, ranges):
        """ranges[r] = (lo, hi) requested by rank r; True if every request's source interval lies in that
        rank's blurred slab (same answer on every rank)"""
        D = self.whd[2]
        for r, (lo, hi) in enumerate(ranges):
            fa, fb = own_range(D, r, self.s.world)
            Va, Vb = max(0, fa - self.ghost), min(D, fb + self.ghost)
            s_lo, s_hi = source_range(lo, hi, D, dims[2]) if dims[2] != D else (lo, hi)
            if s_lo < Va or s_hi > Vb:
                return False
        return True

    def level_frame(self, which, level, dims, lo, hi, all_ranges=None):
        be = self.s.be
        w, hh, d = dims
        D = self.whd[2]
        if all_ranges is not None and self._fits_everywhere(dims, all_ranges):
            if level == 0:
                return Slab(self.slabs[which].planes(lo, hi), lo, d, w)
            return self.s._frame(self.slabs[which], self.whd, dims, lo, hi)
        assert level != 0, "frame ghost depth too small for the finest level (raise frame_ghost)"
        key = (which, level)
        if key not in self._gathered:
            self._gathered = {k: v for k, v in self._gathered.items() if k[1] == level}  # drop older levels
            self._gathered[key] = self._gather(which, dims)
        full = self._gathered[key]
        return Slab(full.planes(lo, hi), lo, d, w)

    def _gather(self, which, dims):
        be = self.s.be
        w, hh, d = dims
        W, Hh, D = self.whd
        world = self.s.world
        # output plane o belongs to the rank in whose own full-resolution range its source interval starts
        delta = f32(D) / f32(d)
        starts = np.floor(np.arange(d, dtype=np.float32) * delta).astype(np.int64)
        bounds = [int(np.searchsorted(starts, own_range(D, r, world)[0], side="left")) for r in range(world)] + [d]
        p_lo, p_hi = bounds[self.s.rank], bounds[self.s.rank + 1]
        n_max = max(bounds[r + 1] - bounds[r] for r in range(world))
        piece = be.zeros(w, hh, max(n_max, 1))
        if p_hi > p_lo:
            s_lo, s_hi = source_range(p_lo, p_hi, D, d)
            assert self.Va <= s_lo and s_hi <= self.Vb, "frame ghost depth smaller than one coarse source interval"
            src = Slab(self.slabs[which].planes(s_lo, s_hi), s_lo, D, W)
            out = Slab(piece[:p_hi - p_lo], p_lo, d, w)
            be.resample(src, self.whd, dims, p_lo, p_lo, p_hi, out=out)
        pieces = self.s.comm.all_gather(piece)
        full = be.empty(w, hh, d)
        for r in range(world):
            n = bounds[r + 1] - bounds[r]
            if n > 0:
                full[bounds[r]:bounds[r + 1]] = pieces[r][:n]
        self.s.stats["frame_gathers"] = self.s.stats.get("frame_gathers", 0) + 1
        return Slab(full, 0, d, w)


# =====================================================================================================
class ShardedFlowSolver:
    """Coarse-to-fine solve of OpticalFlowE::ComputeFlow (optical_flow_e.cpp:132-601) with every large
    level sharded along z over the ranks of `group`."""

    def __init__(self, backend, rank=None, world=None, min_planes_per_rank=12, min_voxels_per_rank=1 << 18,
                 comm=None):
        self.be = backend
        self.comm = comm if comm is not None else TorchComm(rank, world)
        self.rank = self.comm.rank
        self.world = self.comm.world
        self.min_planes = min_planes_per_rank
        self.min_voxels = min_voxels_per_rank
        self.stats = {"sharded_levels": 0, "replicated_levels": 0, "exchanges": 0, "exchange_bytes": 0,
                      "overlapped_levels": 0}
        # Overlap of the per-outer-iteration halo exchange with compute (see _outer_loop_overlapped).
        # Opt-in (FLOW3D_OVERLAP=1): measured on 2 B200s at 512^3 it hides the exchange (118-190 ms -> 36 ms
        # of waits) but the short boundary launches (6..16 planes) run ~3x less efficiently than the
        # interior ones and cost 200 ms + 70 ms of staging copies, a net loss there (2.13 s vs 1.97-2.12 s).
        import os
        env = os.environ.get("FLOW3D_OVERLAP", "")
        self.overlap_min_world = 2 if env == "1" else (1 << 30)
        self.profile = False      # record CUDA events around every batch of sweeps (CabiBackend only)
        self.sweep_events = []    # (start, end, voxel_sweeps, phi_ksi_voxels) per outer iteration
        self.phase_marks = []     # (phase name, CUDA event): time until the next mark belongs to the phase

    def _mark(self, name):
        """profiling only: the device time from here to the next mark is attributed to phase `name`"""
        if self.profile and self.be.name == "cabi":
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.phase_marks.append((name, e))

    def phase_profile(self):
        """{phase: milliseconds} of the recorded marks (device time, this rank); clears the record"""
        out = {}
        m = self.phase_marks
        for (name, e0), (_, e1) in zip(m[:-1], m[1:]):
            if name != "end":
                out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        self.phase_marks = []
        return out

    # ---- neighbour exchange of ghost planes -------------------------------------------------------
    def _exchange(self, fields, A, B, a, b, H, D):
        """fill ghosts [A,a) and [b,B) of every tensor in `fields` from the z neighbours' owned planes"""
        if self.world == 1:
            return
        ops = []
        lo_n, hi_n = self.rank - 1, self.rank + 1
        for t in fields:
            if lo_n >= 0 and a > A:  # my lower ghosts <- rank-1's top planes; rank-1's upper ghosts <- my bottom planes
                ops.append(("recv", t[0:a - A], lo_n))
                ops.append(("send", t[a - A:a - A + min(H, b - a)], lo_n))
            if hi_n < self.world and B > b:
                ops.append(("send", t[b - A - min(H, b - a):b - A], hi_n))
                ops.append(("recv", t[b - A:B - A], hi_n))
        if ops:
            self.comm.sendrecv(ops)
            self.stats["exchanges"] += 1
            self.stats["exchange_bytes"] += sum(t.numel() * 4 for kind, t, _ in ops if kind == "send")

    def _is_sharded(self, dims, H):
        w, h, d = dims
        if self.world == 1:
            return False
        per = d // self.world
        return per >= max(self.min_planes, 2 * H) and w * h * per >= self.min_voxels

    # ---- the solve ---------------------------------------------------------------------------------
    def compute(self, frame_0, frame_1, params=None, level_cb=None, width=None, return_device=False):
        """frame_0/frame_1: FULL volumes given to every rank: numpy (D,H,W) arrays, or backend tensors
        (D,H,ld) already on the device together with `width`.  Returns (a, b, [u,v,w]): this rank's owned
        plane range of the finest level and those planes (numpy, or device tensors (b-a,H,ld) when
        return_device) -- the full volume on every rank if the finest level was too small to shard."""
        be = self.be
        P = dict(DEFAULTS)
        P.update(params or {})
        inner, outer = int(P["inner_iterations_count"]), int(P["outer_iterations_count"])
        H = inner + 1
        if isinstance(frame_0, np.ndarray):
            D, Hh, W = frame_0.shape
            F0 = be.from_numpy_full(frame_0)
            F1 = be.from_numpy_full(frame_1)
        else:
            D, Hh, _ = frame_0.shape
            W = int(width)
            be.w_full = W
            F0, F1 = frame_0, frame_1
        sched = self._schedule(W, Hh, D, P)
        prof = self.profile
        if P["gaussian_sigma"] > 0:
            F0, F1 = be.blur(F0, P["gaussian_sigma"]), be.blur(F1, P["gaussian_sigma"])
        frames = ReplicatedFrames(self, F0, F1, (W, Hh, D))
        return self._solve(frames, (W, Hh, D), sched, P, level_cb, return_device)

    def compute_slabs(self, raw0, raw1, raw_z0, whd, params=None, level_cb=None, return_device=False,
                      frame_ghost=32):
        """Like compute(), but every rank passes only a z-slab of the two RAW frames (backend tensors
        (dl,H,ld) or numpy (dl,H,W)) starting at global plane raw_z0 and covering at least
        ShardedFrames.input_planes(...): nothing full-size is ever held by one rank (2048^3 and up)."""
        be = self.be
        P = dict(DEFAULTS)
        P.update(params or {})
        W, Hh, D = whd
        if isinstance(raw0, np.ndarray):
            raw0, raw1 = be.from_numpy_full(raw0), be.from_numpy_full(raw1)
        be.w_full = W
        sched = self._schedule(W, Hh, D, P)
        frames = ShardedFrames(self, raw0, raw1, raw_z0, (W, Hh, D), P["gaussian_sigma"], frame_ghost)
        return self._solve(frames, (W, Hh, D), sched, P, level_cb, return_device)

    def _solve(self, frames, whd, sched, P, level_cb, return_device):
        be = self.be
        W, Hh, D = whd
        inner, outer = int(P["inner_iterations_count"]), int(P["outer_iterations_count"])
        H = inner + 1
        prof = self.profile
        prev = None  # (dims, [u,v,w] Slabs, valid_lo, valid_hi)
        for (level, dims, h) in sched:
            w, hh, d = dims
            sharded = self._is_sharded(dims, H)
            a, b = own_range(d, self.rank, self.world) if sharded else (0, d)
            A, B = (max(0, a - H), min(d, b + H)) if sharded else (0, d)
            dl = B - A
            self.stats["sharded_levels" if sharded else "replicated_levels"] += 1

            # ---- flow at this level: zeros, or box-prolongation of the previous level (:304-344) -----
            self._mark("prolongation")
            flow = [Slab(be.zeros(w, hh, dl), A, d, w) for _ in range(3)]
            if prev is not None:
                pdims, pflow, pv_lo, pv_hi = prev
                s_lo, s_hi = source_range(a, b, pdims[2], d)
                assert pv_lo <= s_lo and s_hi <= pv_hi, "prolongation needs planes outside the valid range"
                for c in range(3):
                    src = Slab(pflow[c].planes(s_lo, s_hi), s_lo, pdims[2], pdims[0])
                    be.resample(src, pdims, dims, A, a, b, out=flow[c])
                self._mark("flow_ghost_exchange")
                self._exchange([f.t for f in flow], A, B, a, b, H, d)
            self._mark("level_frames")
            # ---- frames of this level, resampled from the replicated full-resolution frames -------------
            hz = h[2]
            wmax = be.absmax(flow[2]) if prev is not None else 0.0
            if frames.global_reach and self.world > 1:  # same reach on every rank => same local/gather decision
                wmax = self.comm.allreduce_max(wmax, be.dev)
            reach = int(math.ceil(wmax / float(hz))) + 2 if prev is not None else 1
            A1, B1 = max(0, A - reach), min(d, B + reach)
            if frames.global_reach:
                rng0, rng1 = [], []
                for r in range(self.world):
                    ra, rb = own_range(d, r, self.world) if sharded else (0, d)
                    rA, rB = (max(0, ra - H), min(d, rb + H)) if sharded else (0, d)
                    rng0.append((rA, rB))
                    rng1.append((max(0, rA - reach), min(d, rB + reach)))
                f0l = frames.level_frame(0, level, dims, A, B, rng0)
                f1l = frames.level_frame(1, level, dims, A1, B1, rng1)
            else:
                f0l = frames.level_frame(0, level, dims, A, B)
                f1l = frames.level_frame(1, level, dims, A1, B1)
            # ---- warp + derivatives on every plane the solver touches ----------------------------------
            lo1 = A if A == 0 else A + 1
            hi1 = B if B == d else B - 1
            self._mark("warp_derivs")
            terms = be.warp_terms(f0l, f1l, flow[0], flow[1], flow[2], h, lo1, hi1)
            # ---- solver (cuda_operation_solve.cpp:183-257) ----------------------------------------------
            self._mark("alloc_zero")
            d_cur = [be.zeros(w, hh, dl) for _ in range(3)]
            d_alt = [be.zeros(w, hh, dl) for _ in range(3)]
            phi, ksi = be.zeros(w, hh, dl), be.zeros(w, hh, dl)
            use_overlap = (sharded and self.world >= self.overlap_min_world and d // self.world >= 4 * H
                           and be.name != "cabi-streamed")
            if use_overlap:
                self.stats["overlapped_levels"] += 1
                d_cur, d_alt = self._outer_loop_overlapped(terms, flow, d_cur, d_alt, phi, ksi, h, P, (A, B, a, b), d)
            else:
                for _ in range(outer):
                    self._mark("solver")
                    if prof:
                        e0 = torch.cuda.Event(enable_timing=True)
                        e0.record()
                    d_cur, d_alt = be.outer_iteration(terms, flow[0], flow[1], flow[2], d_cur, d_alt, phi, ksi, h,
                                                      inner, P["equation_alpha"], P["equation_smoothness"],
                                                      P["equation_data"], lo1, hi1)
                    if prof:
                        e1 = torch.cuda.Event(enable_timing=True)
                        e1.record()
                        self.sweep_events.append((e0, e1, self._sweep_units(w, hh, d, lo1, hi1, inner),
                                                  w * hh * (hi1 - lo1)))
                    if sharded:
                        self._mark("halo_exchange")
                        self._exchange(d_cur, A, B, a, b, H, d)
            self._mark("update")
            # ---- u += du (:420-438), valid on the whole buffer because the last exchange refreshed du ----
            be.add3(flow, d_cur)
            # ---- median (:443-473) on everything whose +-r/2 neighbourhood is valid ---------------------
            r = int(P["median_radius"])
            r2 = (r - 1 if (r % 2 == 0 and r > 1) else r) // 2
            m_lo = A if A == 0 else A + r2
            m_hi = B if B == d else B - r2
            self._mark("median")
            for c in range(3):
                be.median(flow[c], d_alt[c], r, m_lo, m_hi)
                flow[c] = Slab(d_alt[c], A, d, w)
                d_alt[c] = None
            prev = (dims, flow, m_lo, m_hi)
            self._mark("end")
            if level_cb is not None:
                level_cb(level, dims, (a, b), [be.to_numpy(f.planes(a, b), w) for f in flow])
        dims, flow, _, _ = prev
        a, b = own_range(dims[2], self.rank, self.world) if self._is_sharded(dims, H) else (0, dims[2])
        if return_device:
            return a, b, [f.planes(a, b) for f in flow]
        return a, b, [be.to_numpy(f.planes(a, b), dims[0]) for f in flow]

    @staticmethod
    def _sweep_units(w, hh, d, lo1, hi1, inner):
        units = 0
        for j in range(1, inner + 1):
            units += w * hh * ((hi1 if hi1 == d else hi1 - j) - (lo1 if lo1 == 0 else lo1 + j))
        return units

    def _outer_loop_overlapped(self, terms, flow, d_cur, d_alt, phi, ksi, h, P, ranges, d):
        """The outer iterations of one sharded level with the halo exchange hidden behind compute.

        The values an outer iteration produces at plane p depend on the iterate it starts from on
        [p-H, p+H] only (phi: +-1, each of the `inner` sweeps: +-1; H = inner+1).  So the iteration splits
        into an INTERIOR task, run on the owned planes [a,b) alone (final values valid on [a+H, b-H), no
        ghost needed), and one BOUNDARY task per neighbour, run in a private 3H-plane buffer holding the
        neighbour's H ghost planes and the adjacent 2H owned planes (final values valid on the H owned
        planes next to the neighbour).  Per iteration: stage the 2H owned planes into the boundary
        buffers, post the neighbour exchange (sends/receives use the boundary buffers only), run the
        interior task while the planes travel, wait, run the boundary tasks, copy their H final planes
        into the iterate.  Same kernels, same per-voxel arithmetic, same result as the serial scheme;
        ~7 % more voxel-sweeps at 128 planes per rank."""
        be = self.be
        A, B, a, b = ranges
        inner, outer = int(P["inner_iterations_count"]), int(P["outer_iterations_count"])
        H = inner + 1
        w, hh = flow[0].w, flow[0].t.shape[1]
        alpha, eps_s, eps_d = P["equation_alpha"], P["equation_smoothness"], P["equation_data"]
        prof = self.profile
        sides = []  # (neighbour rank, region [rA, rB), ghost planes, owned planes to send, planes to copy back)
        if a > A:
            sides.append({"peer": self.rank - 1, "rA": A, "rB": a + 2 * H, "ghost": (A, a), "send": (a, a + H),
                          "own": (a, a + 2 * H), "final": (a, a + H)})
        if B > b:
            sides.append({"peer": self.rank + 1, "rA": b - 2 * H, "rB": B, "ghost": (b, B), "send": (b - H, b),
                          "own": (b - 2 * H, b), "final": (b - H, b)})
        for sd in sides:
            n = sd["rB"] - sd["rA"]
            sd["cur"] = [be.zeros(w, hh, n) for _ in range(3)]
            sd["alt"] = [be.zeros(w, hh, n) for _ in range(3)]
            sd["phi"], sd["ksi"] = be.zeros(w, hh, n), be.zeros(w, hh, n)
            o = sd["rA"] - A
            sd["terms"] = [t[o:o + n] for t in terms]
            sd["flow"] = [Slab(f.t[o:o + n], sd["rA"], d, w) for f in flow]
            sd["lo1"] = sd["rA"] if sd["rA"] == 0 else sd["rA"] + 1
            sd["hi1"] = sd["rB"] if sd["rB"] == d else sd["rB"] - 1
        lo_int = 0 if a == 0 else a + 1   # interior task: the owned planes are its whole world
        hi_int = d if b == d else b - 1
        for it in range(outer):
            handle = None
            if it > 0:  # the first iteration starts from du = 0 everywhere: ghosts are zeros already
                self._mark("halo_stage")
                ops = []
                for sd in sides:
                    rA = sd["rA"]
                    for c in range(3):
                        sd["cur"][c][sd["own"][0] - rA:sd["own"][1] - rA].copy_(d_cur[c][sd["own"][0] - A:sd["own"][1] - A])
                for c in range(3):  # field-major order on both sides of a link, as in _exchange
                    for sd in sides:
                        rA = sd["rA"]
                        g0, g1 = sd["ghost"]
                        s0, s1 = sd["send"]
                        if sd["peer"] < self.rank:
                            ops.append(("recv", sd["cur"][c][g0 - rA:g1 - rA], sd["peer"]))
                            ops.append(("send", sd["cur"][c][s0 - rA:s1 - rA], sd["peer"]))
                        else:
                            ops.append(("send", sd["cur"][c][s0 - rA:s1 - rA], sd["peer"]))
                            ops.append(("recv", sd["cur"][c][g0 - rA:g1 - rA], sd["peer"]))
                handle = self.comm.start_sendrecv(ops)
                self.stats["exchanges"] += 1
                self.stats["exchange_bytes"] += sum(t.numel() * 4 for kind, t, _ in ops if kind == "send")
            self._mark("solver")
            if prof:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            d_cur, d_alt = be.outer_iteration(terms, flow[0], flow[1], flow[2], d_cur, d_alt, phi, ksi, h, inner, alpha,
                                              eps_s, eps_d, lo_int, hi_int)
            if prof:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                self.sweep_events.append((e0, e1, self._sweep_units(w, hh, d, lo_int, hi_int, inner),
                                          w * hh * (hi_int - lo_int)))
            if handle is not None:
                self._mark("halo_wait")
                self.comm.finish_sendrecv(handle)
            self._mark("solver_boundary")
            for sd in sides:
                if prof:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                sd["cur"], sd["alt"] = be.outer_iteration(sd["terms"], sd["flow"][0], sd["flow"][1], sd["flow"][2],
                                                          sd["cur"], sd["alt"], sd["phi"], sd["ksi"], h, inner, alpha,
                                                          eps_s, eps_d, sd["lo1"], sd["hi1"])
                f0, f1 = sd["final"]
                for c in range(3):
                    d_cur[c][f0 - A:f1 - A].copy_(sd["cur"][c][f0 - sd["rA"]:f1 - sd["rA"]])
                if prof:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    self.sweep_events.append((e0, e1, self._sweep_units(w, hh, d, sd["lo1"], sd["hi1"], inner),
                                              w * hh * (sd["hi1"] - sd["lo1"])))
        # the update, the median and the next prolongation read the iterate's ghost planes
        self._mark("halo_exchange")
        self._exchange(d_cur, A, B, a, b, H, d)
        return d_cur, d_alt

    def sweep_profile(self):
        """(milliseconds, voxel-sweeps, phi/ksi voxel-updates) of the recorded outer iterations (each
        record brackets one phi/ksi launch + `inner` sweep launches); clears the record"""
        ms = sum(r[0].elapsed_time(r[1]) for r in self.sweep_events)
        units = sum(r[2] for r in self.sweep_events)
        phi_units = sum(r[3] for r in self.sweep_events)
        self.sweep_events = []
        return ms, units, phi_units

    def _frame(self, full, full_whd, dims, lo, hi):
        """planes [lo, hi) of a level frame, box-resampled from the replicated full-resolution frame
        (optical_flow_e.cpp:279-299: always from full resolution)"""
        be = self.be
        w, hh, d = dims
        s_lo, s_hi = source_range(lo, hi, full_whd[2], d)
        src = Slab(full.planes(s_lo, s_hi), s_lo, full_whd[2], full_whd[0])
        out = Slab(be.empty(w, hh, hi - lo), lo, d, w)
        return be.resample(src, full_whd, dims, lo, lo, hi, out=out)

    @staticmethod
    def _schedule(W, H, D, P):
        from .api import level_schedule
        return level_schedule(W, H, D, P["warp_scale_factor"], P["warp_levels_count"])
