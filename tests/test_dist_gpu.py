"""z-sharded solve on real GPUs (NCCL, one process per GPU): must equal the single-GPU solve bit for bit.
Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, smooth_volume

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, params, out_dir, slabs):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cuda_flow3d_b200.dist import CabiBackend, ShardedFlowSolver
    f0 = smooth_volume(shape, 21)
    f1 = np.ascontiguousarray(np.roll(f0, (1, -1, 2), axis=(0, 1, 2)))
    solver = ShardedFlowSolver(CabiBackend(rank), min_planes_per_rank=8, min_voxels_per_rank=1)
    if slabs:
        from cuda_flow3d_b200.dist import ShardedFrames
        ghost = 16
        lo, hi = ShardedFrames.input_planes(shape[0], rank, world, 2.0, ghost)
        a, b, flow = solver.compute_slabs(np.ascontiguousarray(f0[lo:hi]), np.ascontiguousarray(f1[lo:hi]), lo,
                                          (shape[2], shape[1], shape[0]), params, frame_ghost=ghost)
    else:
        a, b, flow = solver.compute(f0, f1, params)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), a=a, b=b, u=flow[0], v=flow[1], w=flow[2],
             sharded=solver.stats["sharded_levels"], gathers=solver.stats.get("frame_gathers", 0))
    dist.destroy_process_group()


@pytest.mark.parametrize("slabs", [False, True])
def test_sharded_equals_single_gpu(gpu, tmp_path, slabs):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    shape = (96, 40, 72)
    params = dict(outer_iterations_count=3, inner_iterations_count=5, warp_levels_count=14)
    mp.spawn(_worker, args=(world, _free_port(), shape, params, str(tmp_path), slabs), nprocs=world, join=True)
    f0 = smooth_volume(shape, 21)
    f1 = np.ascontiguousarray(np.roll(f0, (1, -1, 2), axis=(0, 1, 2)))
    d, h, w = shape
    of = gpu.OpticalFlowE()
    of.silent = True
    assert of.Initialize(gpu.DataSize4(w, h, d))
    full = dict(gpu.DEFAULTS)
    full.update(params)
    ref = [np.zeros_like(f0) for _ in range(3)]
    of.ComputeFlow(f0, f1, ref[0], ref[1], ref[2], full)
    of.Destroy()
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        a, b = int(z["a"]), int(z["b"])
        assert int(z["sharded"]) >= 2
        for c, name in enumerate("uvw"):
            assert np.array_equal(z[name], ref[c][a:b])
