"""Convergence diagnostics: the update-norm reduction and the solver's per-outer-
iteration records.  Not part of the reference's path (it never evaluates a residual,
cuda_operation_solve.cpp:194-257), so the check is against numpy in float64: the kernel accumulates
float32 differences in double, tolerance 1e-12 relative (summation order only)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(5, 7, 9), (16, 33, 130), (40, 41, 67)])
def test_update_norm_matches_numpy(gpu, shape):
    rng = np.random.default_rng(3)
    a = [rng.standard_normal(shape).astype(np.float32) for _ in range(3)]
    b = [rng.standard_normal(shape).astype(np.float32) for _ in range(3)]
    s, m = gpu.ops.update_norm(a, b)
    d = [(x - y).astype(np.float64) for x, y in zip(a, b)]  # the fp32 difference, as the kernel forms it
    ref_s = sum(float((t * t).sum()) for t in d)
    ref_m = max(float(np.abs(t).max()) for t in d)
    assert abs(s - ref_s) <= 1e-12 * ref_s
    assert m == ref_m
    # a z sub-range (the multi-GPU form: owned planes only)
    s2, m2 = gpu.ops.update_norm(a, b, z_range=(1, shape[0] - 1))
    ref_s2 = sum(float((t[1:-1] * t[1:-1]).sum()) for t in d)
    assert abs(s2 - ref_s2) <= 1e-12 * ref_s2
    assert m2 == max(float(np.abs(t[1:-1]).max()) for t in d)


def test_update_norm_nan_sticks(gpu):
    a = [np.zeros((4, 8, 8), np.float32) for _ in range(3)]
    b = [np.zeros((4, 8, 8), np.float32) for _ in range(3)]
    a[1][2, 3, 4] = np.nan
    s, m = gpu.ops.update_norm(a, b)
    assert np.isnan(s) and np.isnan(m)


def _solve(gpu, f0, f1, P, diag=None, tol=0.0):
    D, H, W = f0.shape
    s = gpu.OpticalFlowE()
    s.silent = True
    assert s.Initialize(gpu.DataSize4(W, H, D))
    if diag:
        s.set_diagnostics(True, tol)
    out = [np.zeros_like(f0) for _ in range(3)]
    s.ComputeFlow(f0, f1, out[0], out[1], out[2], P)
    rec = s.diagnostics() if diag else None
    s.Destroy()
    return out, rec


def test_solver_records_do_not_change_the_result(gpu):
    f0, f1, _ = gpu.ops.synth_pair(48, 40, 36, truth=False)
    P = dict(gpu.DEFAULTS, warp_levels_count=6, outer_iterations_count=5, warp_scale_factor=0.8)
    plain, _ = _solve(gpu, f0, f1, P)
    watched, rec = _solve(gpu, f0, f1, P, diag=True)
    for a, b in zip(plain, watched):
        assert np.array_equal(a, b)  # diagnostics are read-only
    sched = gpu.level_schedule(48, 40, 36, 0.8, 6)
    assert len(rec) == len(sched)
    for n, rms, mx in rec:
        assert n == 5 and len(rms) == 5
        assert all(np.isfinite(rms)) and all(np.isfinite(mx))
        assert all(r <= m + 1e-12 for r, m in zip(rms, mx))
        assert all(r > 0 for r in rms)


def test_update_tolerance_stops_early(gpu):
    f0, f1, _ = gpu.ops.synth_pair(48, 40, 36, truth=False)
    P = dict(gpu.DEFAULTS, warp_levels_count=4, outer_iterations_count=12, warp_scale_factor=0.8)
    _, full = _solve(gpu, f0, f1, P, diag=True)
    tol = float(np.median([r[1][3] for r in full]))  # a value most levels reach around iteration 4
    _, early = _solve(gpu, f0, f1, P, diag=True, tol=tol)
    assert any(n < 12 for n, _, _ in early)
    for n, rms, _ in early:
        assert n == len(rms)
        if n < 12:
            assert rms[-1] < tol and all(r >= tol for r in rms[:-1])
