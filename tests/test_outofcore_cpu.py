"""Out-of-core orchestration on the CPU: the virtual-rank threads, the in-process
communicator and the slab bookkeeping, with the test oracle's slab functions as the compute backend.
The streamed result must equal the single-solve oracle result bit for bit."""
import numpy as np

from cuda_flow3d_b200.dist import ShardedFlowSolver
from oracle_backend import OracleBackend
from cuda_flow3d_b200.outofcore import OutOfCoreFlowSolver

PARAMS = dict(warp_levels_count=4, warp_scale_factor=0.8, outer_iterations_count=2, inner_iterations_count=2,
              median_radius=3, gaussian_sigma=1.0)


def _pair(shape, seed=5):
    rng = np.random.default_rng(seed)
    base = rng.random(shape).astype(np.float32) * 255
    for ax in range(3):  # smooth a little so the flow is not pure noise
        base = (base + np.roll(base, 1, ax) + np.roll(base, -1, ax)) / np.float32(3)
    return base.astype(np.float32), np.roll(base, 1, axis=2).astype(np.float32)


def _single(oracle, f0, f1):
    solver = ShardedFlowSolver(OracleBackend(oracle), rank=0, world=1)
    a, b, flow = solver.compute(f0, f1, PARAMS)
    assert (a, b) == (0, f0.shape[0])
    return flow


def test_three_virtual_slabs_equal_single_solve(oracle):
    f0, f1 = _pair((40, 14, 18))
    want = _single(oracle, f0, f1)
    ooc = OutOfCoreFlowSolver(slabs=3, backend_factory=lambda k: OracleBackend(oracle), frame_ghost=8,
                              min_planes_per_slab=6, min_voxels_per_slab=1)
    got = ooc.compute(f0, f1, PARAMS)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w)


def test_volume_too_thin_to_cut_runs_replicated(oracle):
    f0, f1 = _pair((9, 12, 16))
    want = _single(oracle, f0, f1)
    got = OutOfCoreFlowSolver(slabs=2, backend_factory=lambda k: OracleBackend(oracle), frame_ghost=8).compute(f0, f1, PARAMS)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_a_failing_slab_raises_instead_of_hanging(oracle):
    f0, f1 = _pair((40, 14, 18))

    class Boom(OracleBackend):
        def median(self, *a, **k):
            raise ValueError("boom")

    ooc = OutOfCoreFlowSolver(slabs=3, backend_factory=lambda k: Boom(oracle) if k == 1 else OracleBackend(oracle),
                              frame_ghost=8, min_planes_per_slab=6, min_voxels_per_slab=1)
    try:
        ooc.compute(f0, f1, PARAMS)
    except ValueError as e:
        assert "boom" in str(e)
    else:
        raise AssertionError("expected the slab's error to propagate")
