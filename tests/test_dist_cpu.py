"""Host logic of the z-sharded multi-GPU solve (cuda_flow3d_b200/dist.py) on CPU: world_size-2 gloo
processes run the same orchestration as the GPU path with the test oracle's slab functions as the compute
backend, and must reproduce the single-process oracle solve BIT FOR BIT (Jacobi => sharding is exact)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, smooth_volume


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, params, out_dir, slabs=False, overlap=None):
    import sys
    if overlap is not None:
        os.environ["FLOW3D_OVERLAP"] = "1" if overlap else "0"
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    from cuda_flow3d_b200.dist import ShardedFlowSolver
    from oracle_backend import OracleBackend
    o = Oracle()
    o.set_num_threads(2)
    f0 = smooth_volume(shape, 21)
    f1 = np.ascontiguousarray(np.roll(f0, (1, -1, 2), axis=(0, 1, 2)))
    solver = ShardedFlowSolver(OracleBackend(o), min_planes_per_rank=8, min_voxels_per_rank=1)
    if slabs:
        # every rank is given only the planes of the raw frames it needs (sharded frames)
        from cuda_flow3d_b200.dist import ShardedFrames
        ghost = 12
        lo, hi = ShardedFrames.input_planes(shape[0], rank, world, params.get("gaussian_sigma", 2.0), ghost)
        a, b, flow = solver.compute_slabs(np.ascontiguousarray(f0[lo:hi]), np.ascontiguousarray(f1[lo:hi]), lo,
                                          (shape[2], shape[1], shape[0]), params, frame_ghost=ghost)
    else:
        a, b, flow = solver.compute(f0, f1, params)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), a=a, b=b, u=flow[0], v=flow[1], w=flow[2],
             sharded=solver.stats["sharded_levels"], exchanges=solver.stats["exchanges"],
             gathers=solver.stats.get("frame_gathers", 0), overlapped=solver.stats["overlapped_levels"])
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,params", [
    ((48, 18, 22), dict(outer_iterations_count=2, inner_iterations_count=3, warp_levels_count=12, median_radius=5)),
    ((36, 16, 20), dict(outer_iterations_count=2, inner_iterations_count=2, warp_levels_count=6, median_radius=3,
                        gaussian_sigma=0.0)),
])
@pytest.mark.parametrize("slabs", [False, True])
def test_two_rank_sharded_solve_equals_single_process(oracle, tmp_path, shape, params, slabs):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, shape, params, str(tmp_path), slabs), nprocs=world, join=True)
    f0 = smooth_volume(shape, 21)
    f1 = np.ascontiguousarray(np.roll(f0, (1, -1, 2), axis=(0, 1, 2)))
    ref = oracle.compute_flow(f0, f1, params)
    covered = 0
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        a, b = int(z["a"]), int(z["b"])
        assert int(z["sharded"]) >= 2, "the test must exercise sharded levels"
        assert int(z["exchanges"]) > 0
        if slabs and params["warp_levels_count"] >= 12:
            assert int(z["gathers"]) > 0, "coarse levels must have used the all-gather path"
        for c, name in enumerate("uvw"):
            assert np.array_equal(z[name], ref[c][a:b]), "rank %d flow_%s planes [%d,%d) differ" % (r, name, a, b)
        covered += b - a
    assert covered == shape[0]


@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("world,shape,params,slabs", [
    (3, (54, 14, 16), dict(outer_iterations_count=3, inner_iterations_count=2, warp_levels_count=5, median_radius=3,
                           warp_scale_factor=0.9), False),
    (4, (64, 12, 14), dict(outer_iterations_count=2, inner_iterations_count=2, warp_levels_count=4, median_radius=5,
                           warp_scale_factor=0.9, gaussian_sigma=1.0), True),
])
def test_more_ranks_with_and_without_overlapped_exchange(oracle, tmp_path, world, shape, params, slabs, overlap):
    """world 3 and 4 (interior ranks with two neighbours, edge ranks with one): the default serial exchange
    and the opt-in boundary-first / interior split of the outer iteration
    (ShardedFlowSolver._outer_loop_overlapped, FLOW3D_OVERLAP=1) both reproduce the single-process solve"""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, shape, params, str(tmp_path), slabs, overlap), nprocs=world, join=True)
    f0 = smooth_volume(shape, 21)
    f1 = np.ascontiguousarray(np.roll(f0, (1, -1, 2), axis=(0, 1, 2)))
    ref = oracle.compute_flow(f0, f1, params)
    covered = 0
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        a, b = int(z["a"]), int(z["b"])
        assert int(z["sharded"]) >= 1
        assert (int(z["overlapped"]) >= 1) == overlap, "overlap flag not honoured"
        for c, name in enumerate("uvw"):
            assert np.array_equal(z[name], ref[c][a:b]), "rank %d flow_%s planes [%d,%d) differ" % (r, name, a, b)
        covered += b - a
    assert covered == shape[0]


def test_partition_and_source_ranges():
    from cuda_flow3d_b200.dist import own_range, source_range
    for d in (5, 17, 128, 1000):
        for world in (1, 2, 4, 8):
            edges = [own_range(d, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == d
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    # the source interval must cover exactly the taps the kernels read
    for (a, b) in [(128, 122), (20, 24), (70, 512), (512, 70)]:
        delta = np.float32(a) / np.float32(b)
        for lo, hi in [(0, b), (3, 7), (b - 2, b)]:
            s_lo, s_hi = source_range(lo, hi, a, b)
            taps = []
            for o in range(lo, hi):
                li = int(np.floor(np.float32(o) * delta))
                ri = int(min(np.float32(a), np.ceil(np.float32(o + 1) * delta)))
                taps += list(range(li, ri))
            assert s_lo <= min(taps) and max(taps) < s_hi
