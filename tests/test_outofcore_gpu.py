"""Host-resident (out-of-core) solve on the GPU: z-slabs streamed through the device by virtual-rank
threads must reproduce the in-core solve bit for bit (the in-core solve is itself pinned to the reference
CUDA build, tests/test_golden.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _incore(gpu, f0, f1, P):
    D, H, W = f0.shape
    s = gpu.OpticalFlowE()
    s.silent = True
    assert s.Initialize(gpu.DataSize4(W, H, D))
    out = [np.zeros_like(f0) for _ in range(3)]
    s.ComputeFlow(f0, f1, out[0], out[1], out[2], P)
    s.Destroy()
    return out


@pytest.mark.parametrize("slabs,concurrency", [(3, 2), (2, 1)])
def test_streamed_slabs_equal_incore_solve(gpu, slabs, concurrency):
    from cuda_flow3d_b200.outofcore import OutOfCoreFlowSolver
    f0, f1, _ = gpu.ops.synth_pair(52, 44, 72, truth=False)
    P = dict(gpu.DEFAULTS, warp_levels_count=8, outer_iterations_count=3, warp_scale_factor=0.9)
    want = _incore(gpu, f0, f1, P)
    ooc = OutOfCoreFlowSolver(device=0, slabs=slabs, concurrency=concurrency, frame_ghost=16, min_voxels_per_slab=1)
    got = ooc.compute(f0, f1, P)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert ooc.stats["h2d_bytes"] > 0 and ooc.stats["d2h_bytes"] > 0


def test_default_parameters_small_volume(gpu):
    """default pyramid (levels limited by the volume), median 5, sigma 2"""
    from cuda_flow3d_b200.outofcore import OutOfCoreFlowSolver
    f0, f1, _ = gpu.ops.synth_pair(40, 36, 64, truth=False)
    P = dict(gpu.DEFAULTS, outer_iterations_count=2)
    want = _incore(gpu, f0, f1, P)
    got = OutOfCoreFlowSolver(slabs=2, frame_ghost=20, min_voxels_per_slab=1).compute(f0, f1, P)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
