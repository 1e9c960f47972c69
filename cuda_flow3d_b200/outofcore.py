"""Host-resident (out-of-core) flow solve: volumes larger than one GPU's memory on ONE GPU.

SURVEY.md 8f rank 2: the reference's second solver, OpticalFlowP (src/optical_flow/optical_flow_p.cpp with
src/cuda_operations/partial_data/*), keeps the volumes in host memory and streams z-slabs with a
one-plane mirrored halo through the device (cuda_operation_solve_p.cpp:358-417); it is disabled in the
reference's main.cpp and does not build with current CUDA.  This is its replacement, built from the
machinery of the multi-GPU solve instead of a second set of kernels:

* the volume is cut into `slabs` z-slabs exactly as ShardedFlowSolver cuts it over ranks; each slab is
  a *virtual rank* running the unmodified sharded algorithm (dist.py) in its own host thread;
* a virtual rank's fields live in (pinned) host memory; every stage call uploads its operands to the
  device, runs the same C-ABI slab kernels and downloads the results (HostStreamedBackend), so the
  device holds the working set of at most `concurrency` stage calls at any time;
* the neighbour exchange, the scalar max and the frame all-gather happen between threads in host
  memory (dist.LocalComm) -- no NCCL, no second GPU;
* the threads use separate CUDA streams, so one slab's PCIe transfers overlap another slab's kernels.

Because each virtual rank executes the same kernels on the same planes as a real rank would, the
result is bit-identical to the in-core solve (tests/test_outofcore_gpu.py), as the multi-GPU result is.
It is PCIe-bound by construction (each outer iteration moves a slab's fields both ways); the in-core
and multi-GPU solvers are the fast paths.
"""
import threading

import numpy as np
import torch

from .api import DEFAULTS
from .dist import CabiBackend, LocalComm, ShardedFlowSolver, ShardedFrames, Slab


class HostStreamedBackend:
    """CabiBackend stage calls on HOST tensors: upload operands, run, download results."""
    name = "cabi-streamed"

    def __init__(self, device, gate, pinned=True, cache_static=True):
        # cache_static: keep a level's read-only solver inputs (fx,fy,fz,ft,u,v,w of this slab) on the
        # device across its outer iterations instead of re-uploading them 40 times (7 of the 19 in-core
        # volumes stay resident; switch off when even that does not fit)
        self.cache_static = cache_static
        self._static_key = None
        self._static_dev = None
        self.inner = CabiBackend(device)
        self.dev = torch.device("cpu")
        self.cuda = self.inner.dev
        self.gate = gate          # semaphore bounding how many stage calls hold device memory at once
        self.pinned = pinned
        self.stream = torch.cuda.Stream(device=self.cuda)
        self.bytes_h2d = 0
        self.bytes_d2h = 0
        self.w_full = None

    # ---- host allocation ---------------------------------------------------------------------------
    def ld(self, w):
        return self.inner.ld(w)

    def _alloc(self, shape, zero):
        try:
            t = torch.empty(shape, dtype=torch.float32, pin_memory=self.pinned)
        except RuntimeError:  # pinned pool exhausted: pageable memory still works, only slower
            t = torch.empty(shape, dtype=torch.float32)
        if zero:
            t.zero_()
        return t

    def empty(self, w, h, dl):
        return self._alloc((dl, h, self.ld(w)), False)

    def zeros(self, w, h, dl):
        return self._alloc((dl, h, self.ld(w)), True)

    # ---- transfer helpers --------------------------------------------------------------------------
    def _up(self, t):
        self.bytes_h2d += t.numel() * 4
        return t.to(self.cuda, non_blocking=True)

    def _up_slab(self, s):
        return Slab(self._up(s.t), s.A, s.dg, s.w)

    def _down(self, dev_t, host_t):
        self.bytes_d2h += host_t.numel() * 4
        host_t.copy_(dev_t, non_blocking=True)

    class _Call:
        """with be._call(): ... -- one stage call: bounded device residency, own stream, synchronous end"""

        def __init__(self, be):
            self.be = be

        def __enter__(self):
            self.be.gate.acquire()
            self.ctx = torch.cuda.stream(self.be.stream)
            self.ctx.__enter__()
            self.be.inner.w_full = self.be.w_full
            return self

        def __exit__(self, *exc):
            try:
                self.be.stream.synchronize()
                self.ctx.__exit__(*exc)
            finally:
                self.be.gate.release()
            return False

    def _call(self):
        return HostStreamedBackend._Call(self)

    # ---- stage calls (same signatures as CabiBackend) ----------------------------------------------
    def blur_slab(self, raw, sigma, lo, hi):
        with self._call():
            out = self.inner.blur_slab(self._up_slab(raw), sigma, lo, hi)
            host = self._alloc(tuple(raw.t.shape), False)
            self._down(out.t, host)
        return Slab(host, raw.A, raw.dg, raw.w)

    def blur(self, full, sigma):
        with self._call():
            out = self.inner.blur(self._up(full), sigma)
            host = self._alloc(tuple(full.shape), False)
            self._down(out, host)
        return host

    def resample(self, src, src_whd_global, out_whd_global, out_A, out_lo, out_hi, out=None):
        with self._call():
            d_out = self._up_slab(out)  # planes outside [out_lo, out_hi) keep their values
            self.inner.resample(self._up_slab(src), src_whd_global, out_whd_global, out_A, out_lo, out_hi, out=d_out)
            self._down(d_out.t, out.t)
        return out

    def warp_terms(self, f0, f1, u, v, w, h, lo, hi):
        with self._call():
            terms = self.inner.warp_terms(self._up_slab(f0), self._up_slab(f1), self._up_slab(u), self._up_slab(v),
                                          self._up_slab(w), h, lo, hi)
            host = [self._alloc(tuple(t.shape), False) for t in terms]
            for d, hst in zip(terms, host):
                self._down(d, hst)
        return host

    def outer_iteration(self, terms, u, v, w, d_cur, d_alt, phi, ksi, h, inner, alpha, eps_s, eps_d, lo1, hi1):
        # Up: the image terms, the flow and the current iterate.  phi/ksi are recomputed at the start of
        # every outer iteration and the alternate buffer is pure scratch (each sweep reads only planes
        # the previous sweep wrote), so those five fields exist on the device only.  Down: the iterate.
        with self._call():
            key = tuple(t.data_ptr() for t in terms) + (u.t.data_ptr(), v.t.data_ptr(), w.t.data_ptr())
            if self.cache_static and self._static_key == key:
                dt, du, dv, dw = self._static_dev
            else:
                self._static_key = self._static_dev = None  # frees the previous level's copies
                dt = [self._up(t) for t in terms]
                du, dv, dw = self._up_slab(u), self._up_slab(v), self._up_slab(w)
                if self.cache_static:
                    self._static_key, self._static_dev = key, (dt, du, dv, dw)
            dc = [self._up(t) for t in d_cur]
            # zero-initialised: with an odd sweep count the result lives in the scratch set, whose outermost
            # ghost planes no sweep writes -- they go back to the host and must not carry garbage / NaN
            da = [torch.zeros_like(t) for t in dc]
            dphi, dksi = torch.zeros_like(dc[0]), torch.zeros_like(dc[0])
            rc, _ = self.inner.outer_iteration(dt, du, dv, dw, dc, da, dphi, dksi, h, inner, alpha, eps_s, eps_d,
                                               lo1, hi1)
            for d, hst in zip(rc, d_cur):
                self._down(d, hst)
        return d_cur, d_alt

    def add3(self, flow, d):
        self._static_key = self._static_dev = None  # the level's solver loop is over
        with self._call():
            df = [self._up_slab(f) for f in flow]
            self.inner.add3(df, [self._up(t) for t in d])
            for x, f in zip(df, flow):
                self._down(x.t, f.t)

    def median(self, src, dst_t, radius, lo, hi):
        with self._call():
            d_src = self._up_slab(src)
            d_dst = torch.empty_like(d_src.t)  # planes outside [lo, hi) are never read back (see dist.py)
            self.inner.median(d_src, d_dst, radius, lo, hi)
            self._down(d_dst, dst_t)

    def absmax(self, s):
        return float(s.t[:, :, :s.w].abs().max()) if s.t.numel() else 0.0

    def from_numpy_full(self, a):
        d, h, w = a.shape
        t = self.zeros(w, h, d)
        t[:, :, :w] = torch.from_numpy(np.ascontiguousarray(a, np.float32))
        self.w_full = w
        return t

    def to_numpy(self, t, w):
        return t[:, :, :w].contiguous().numpy().copy()


class OutOfCoreFlowSolver:
    """Same call shape as OpticalFlowE.ComputeFlow on numpy volumes, for volumes whose 19-volume arena
    does not fit the device: `slabs` virtual ranks streamed through one GPU."""

    def __init__(self, device=0, slabs=4, concurrency=2, pinned=True, backend_factory=None, frame_ghost=32,
                 min_planes_per_slab=12, min_voxels_per_slab=1 << 18, cache_static=True):
        self.cache_static = cache_static
        self.device, self.slabs, self.concurrency, self.pinned = int(device), int(slabs), int(concurrency), pinned
        self.backend_factory = backend_factory  # tests inject a CPU backend per virtual rank
        self.frame_ghost = frame_ghost
        self.min_planes, self.min_voxels = min_planes_per_slab, min_voxels_per_slab
        self.stats = {}

    def compute(self, frame_0, frame_1, params=None):
        """frame_0 / frame_1: numpy float32 (D,H,W).  Returns [u, v, w] numpy (D,H,W)."""
        P = dict(DEFAULTS)
        P.update(params or {})
        D, Hh, W = frame_0.shape
        K = self.slabs
        shared = LocalComm.Shared(K)
        gate = threading.Semaphore(max(1, self.concurrency))
        results, errors, backends = [None] * K, [None] * K, [None] * K

        def work(k):
            try:
                be = self.backend_factory(k) if self.backend_factory else HostStreamedBackend(self.device, gate, self.pinned, self.cache_static)
                backends[k] = be
                solver = ShardedFlowSolver(be, comm=LocalComm(shared, k), min_planes_per_rank=self.min_planes,
                                           min_voxels_per_rank=self.min_voxels)
                z_lo, z_hi = ShardedFrames.input_planes(D, k, K, P["gaussian_sigma"], self.frame_ghost)
                raw0 = be.from_numpy_full(frame_0[z_lo:z_hi])
                raw1 = be.from_numpy_full(frame_1[z_lo:z_hi])
                results[k] = solver.compute_slabs(raw0, raw1, z_lo, (W, Hh, D), P, frame_ghost=self.frame_ghost)
            except BaseException as e:  # noqa: BLE001 -- re-raised in the caller's thread
                errors[k] = e
                shared.failed.set()
                shared.barrier.abort()

        threads = [threading.Thread(target=work, args=(k,), name="flow3d-slab-%d" % k) for k in range(K)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        secondary = (threading.BrokenBarrierError, RuntimeError)
        for e in sorted((e for e in errors if e is not None), key=lambda e: isinstance(e, secondary)):
            raise e  # the root cause first, not a peer's "barrier broken"
        out = [np.empty((D, Hh, W), np.float32) for _ in range(3)]
        covered = 0
        for a, b, flow in results:
            for c in range(3):
                out[c][a:b] = flow[c]
            covered = max(covered, b)
        if results[0][0] == 0 and results[0][1] == D:  # finest level too small to shard: every slab holds it all
            covered = D
        assert covered == D
        self.stats = {"h2d_bytes": sum(getattr(b, "bytes_h2d", 0) for b in backends),
                      "d2h_bytes": sum(getattr(b, "bytes_d2h", 0) for b in backends), "slabs": K}
        return out
