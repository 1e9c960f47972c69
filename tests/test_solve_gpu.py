"""Full pyramid solve on the GPU (through the reference-shaped OpticalFlowE mirror -> C ABI) against
the CPU oracle: bit-exact per level and at the end, on small volumes the oracle finishes in seconds."""
import numpy as np
import pytest

from conftest import smooth_volume

pytestmark = pytest.mark.gpu


def _solve(gpu, f0, f1, params, per_level=None):
    d, h, w = f0.shape
    of = gpu.OpticalFlowE()
    of.silent = True
    assert of.Initialize(gpu.DataSize4(w, h, d))
    if per_level is not None:
        of.set_level_callback(lambda lv, dims, u, v, ww: per_level.append((lv, dims, u, v, ww)))
    p = gpu.OperationParameters()
    vals = dict(gpu.DEFAULTS)
    vals.update(params)
    for k, v in vals.items():
        p.PushValuePtr(k, v)
    out = [np.zeros_like(f0) for _ in range(3)]
    of.ComputeFlow(f0, f1, out[0], out[1], out[2], p)
    assert of.last_status == 0
    of.Destroy()
    return out


@pytest.mark.parametrize("shape,params", [
    ((24, 28, 36), dict(outer_iterations_count=3, inner_iterations_count=5, warp_levels_count=8)),
    ((5, 40, 70), dict(outer_iterations_count=2, inner_iterations_count=3, warp_levels_count=40, median_radius=3)),
    ((20, 20, 20), dict(outer_iterations_count=2, inner_iterations_count=5, warp_levels_count=40, gaussian_sigma=0.0,
                        median_radius=1)),
])
def test_compute_flow_matches_oracle(gpu, oracle, shape, params):
    f0 = smooth_volume(shape, 21)
    f1 = np.ascontiguousarray(np.roll(f0, (1, -1, 2), axis=(0, 1, 2)))
    levels_gpu, levels_cpu = [], []
    got = _solve(gpu, f0, f1, params, levels_gpu)
    ref = oracle.compute_flow(f0, f1, params, lambda lv, dims, u, v, w: levels_cpu.append((lv, dims, u, v, w)))
    assert [x[:2] for x in levels_gpu] == [x[:2] for x in levels_cpu]
    for a, b in zip(levels_gpu, levels_cpu):
        for i in (2, 3, 4):
            assert np.array_equal(a[i], b[i]), "level %d differs" % a[0]
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


def test_missing_parameter_returns_early(gpu, capsys):
    of = gpu.OpticalFlowE()
    of.silent = True
    assert of.Initialize(gpu.DataSize4(16, 16, 16))
    p = gpu.OperationParameters()
    p.PushValuePtr("warp_levels_count", 3)
    z = np.zeros((16, 16, 16), np.float32)
    of.ComputeFlow(z, z, z.copy(), z.copy(), z.copy(), p)
    assert "Missing parameter" in capsys.readouterr().out
    of.Destroy()


def test_uninitialized_solver_is_refused(gpu, capsys):
    of = gpu.OpticalFlowE()
    z = np.zeros((8, 8, 8), np.float32)
    of.ComputeFlow(z, z, z.copy(), z.copy(), z.copy(), gpu.OperationParameters())
    assert "was not initialized" in capsys.readouterr().out


def test_synthetic_known_displacement(gpu):
    """64^3 analytic pair (SURVEY 8d config-3 generator): the solve must recover the rigid motion far
    better than zero flow does."""
    W = 64
    f0, f1, truth = gpu.ops.synth_pair(W, W, W)
    got = _solve(gpu, f0, f1, dict(outer_iterations_count=10))
    epe = np.sqrt(sum((g - t) ** 2 for g, t in zip(got, truth)))
    epe0 = np.sqrt(sum(t ** 2 for t in truth))
    inner = (slice(12, -12),) * 3
    assert epe[inner].mean() < 0.25 * epe0[inner].mean()
