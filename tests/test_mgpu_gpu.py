"""The C++ z-sharded solver (libflow3d_b200_mgpu.so: C-ABI slab stages + NCCL send/recv) must equal the
single-GPU solve BIT FOR BIT (point-Jacobi: no order dependence).

* world = 1 runs on any box: the C++ orchestration (slab frames, blur with halo, prolongation, warp reach,
  outer iterations, median) against OpticalFlowE.
* world >= 2 needs that many GPUs (`gpurun --gpus 2`): ranks as THREADS of one process
  (flow3d_mgpu_compute_host, what OpticalFlowE::SetDevices and `flow3d_cli --gpus` use) and ranks as
  PROCESSES (what bench.py uses under torchrun), with the thresholds lowered so that small levels shard,
  frames are all-gathered on coarse levels and interior ranks (two neighbours) exist at world 4.
"""
import ctypes as C
import os
import time

import numpy as np
import pytest

from conftest import ROOT, smooth_volume

pytestmark = pytest.mark.gpu


def _pair(shape):
    f0 = smooth_volume(shape, 21)
    f1 = np.ascontiguousarray(np.roll(f0, (1, -1, 2), axis=(0, 1, 2)))
    return f0, f1


def _single(gpu, f0, f1, params):
    d, h, w = f0.shape
    of = gpu.OpticalFlowE()
    of.silent = True
    assert of.Initialize(gpu.DataSize4(w, h, d))
    full = dict(gpu.DEFAULTS)
    full.update(params)
    ref = [np.zeros_like(f0) for _ in range(3)]
    of.ComputeFlow(f0, f1, ref[0], ref[1], ref[2], full)
    of.Destroy()
    return ref


def _rank_solve(gpu, rank, world, uid, f0, f1, params, device, min_planes=8, min_voxels=1, ghost=16):
    """one rank: upload its raw slab, solve, download its planes; returns (a, b, [u,v,w], stats)"""
    import cuda_flow3d_b200.mgpu as m
    from cuda_flow3d_b200._lib import check, load
    d, h, w = f0.shape
    check(load().flow3d_set_device(device), "set_device")
    s = m.ShardedSolver(w, h, d, device, rank, world, uid)
    s.set_thresholds(min_planes, min_voxels)
    p = gpu.api.make_params(params)
    lo, hi = m.input_planes(d, rank, world, p.gaussian_sigma, ghost)
    r0 = gpu.DeviceVolume.from_numpy(f0[lo:hi])
    r1 = gpu.DeviceVolume.from_numpy(f1[lo:hi])
    pa, pb = s.output_planes(p)
    outs = [gpu.DeviceVolume.zeros((w, h, pb - pa)) for _ in range(3)]
    a, b = s.compute(r0.ptr.value, r1.ptr.value, lo, hi - lo, r0.ld, p, ghost, [o.ptr.value for o in outs], pb - pa)
    assert (a, b) == (pa, pb)
    flows = [o.numpy() for o in outs]
    st = s.stats()
    s.destroy()
    return a, b, flows, st


PARAMS = dict(outer_iterations_count=3, inner_iterations_count=5, warp_levels_count=14)


@pytest.mark.parametrize("shape", [(40, 36, 52), (5, 44, 60)])
def test_world1_equals_single_gpu(gpu, shape):
    f0, f1 = _pair(shape)
    ref = _single(gpu, f0, f1, PARAMS)
    a, b, flows, st = _rank_solve(gpu, 0, 1, None, f0, f1, PARAMS, 0)
    assert (a, b) == (0, shape[0]) and st["sharded_levels"] == 0
    for c in range(3):
        assert np.array_equal(flows[c], ref[c])


def _n_gpus(gpu):
    return gpu.load().flow3d_device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_threads_sharded_equals_single_gpu(gpu, world, monkeypatch):
    if _n_gpus(gpu) < world:
        pytest.skip("needs >= %d GPUs" % world)
    import cuda_flow3d_b200.mgpu as m
    monkeypatch.setenv("FLOW3D_MGPU_MIN_PLANES", "8")
    monkeypatch.setenv("FLOW3D_MGPU_MIN_VOXELS", "1")
    shape = (48 * world, 40, 72)
    f0, f1 = _pair(shape)
    ref = _single(gpu, f0, f1, PARAMS)
    flows, ms = m.compute_host(f0, f1, list(range(world)), PARAMS)
    assert ms > 0
    for c in range(3):
        assert np.array_equal(flows[c], ref[c])
    # persistent ranks: a second solve on the same communicator
    flows2, _ = m.compute_host(f0, f1, list(range(world)), PARAMS, persistent=True)
    flows3, _ = m.compute_host(f0, f1, list(range(world)), PARAMS, persistent=True)
    m.release_host_group()
    for c in range(3):
        assert np.array_equal(flows2[c], ref[c]) and np.array_equal(flows3[c], ref[c])


def _proc_worker(rank, world, shape, params, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cuda_flow3d_b200 as gpu
    import cuda_flow3d_b200.mgpu as m
    idfile = os.path.join(out_dir, "nccl_id.bin")
    if rank == 0:
        uid = m.unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise RuntimeError("no NCCL id from rank 0")
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    f0, f1 = _pair(shape)
    a, b, flows, st = _rank_solve(gpu, rank, world, uid, f0, f1, params, rank)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), a=a, b=b, u=flows[0], v=flows[1], w=flows[2],
             sharded=st["sharded_levels"], gathers=st["frame_gathers"], exchanges=st["exchanges"])


@pytest.mark.parametrize("world", [2, 4])
def test_processes_sharded_equals_single_gpu(gpu, tmp_path, world):
    if _n_gpus(gpu) < world:
        pytest.skip("needs >= %d GPUs" % world)
    import multiprocessing as mp
    shape = (48 * world, 40, 72)
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_proc_worker, args=(r, world, shape, PARAMS, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    f0, f1 = _pair(shape)
    ref = _single(gpu, f0, f1, PARAMS)
    covered = 0
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        a, b = int(z["a"]), int(z["b"])
        covered += b - a
        assert int(z["sharded"]) >= 2 and int(z["exchanges"]) > 0
        for c, name in enumerate("uvw"):
            assert np.array_equal(z[name], ref[c][a:b])
    assert covered == shape[0]
