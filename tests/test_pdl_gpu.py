"""Programmatic dependent launch of a level's phi / sweep chain (include/flow3d_c.h: flow3d_set_pdl,
csrc/common.cuh: launch_chain_kernel): the kernels and their arguments are the ordinary ones, only the
launch attribute differs, so the flows must be BIT-IDENTICAL with the chain on and off -- on tiny levels
(many launches resident at once), ragged ones and a thin slab -- and equal to the CPU oracle."""
import numpy as np
import pytest

from conftest import smooth_volume

pytestmark = pytest.mark.gpu


def _solve(gpu, f0, f1, params):
    d, h, w = f0.shape
    of = gpu.OpticalFlowE()
    of.silent = True
    assert of.Initialize(gpu.DataSize4(w, h, d))
    full = dict(gpu.DEFAULTS)
    full.update(params)
    out = [np.zeros_like(f0) for _ in range(3)]
    of.ComputeFlow(f0, f1, out[0], out[1], out[2], full)
    assert of.last_status == 0
    of.Destroy()
    return out


@pytest.mark.parametrize("shape,params", [
    ((18, 20, 22), dict(outer_iterations_count=40, inner_iterations_count=5, warp_levels_count=40)),
    ((40, 36, 52), dict(outer_iterations_count=12, inner_iterations_count=5, warp_levels_count=14)),
    ((5, 44, 60), dict(outer_iterations_count=8, inner_iterations_count=4, warp_levels_count=40)),
    ((33, 70, 129), dict(outer_iterations_count=6, inner_iterations_count=5, warp_levels_count=6)),
])
def test_chain_on_equals_chain_off(gpu, shape, params):
    L = gpu.load()
    f0 = smooth_volume(shape, 5)
    f1 = np.ascontiguousarray(np.roll(f0, (1, 2, -1), axis=(0, 1, 2)))
    try:
        L.flow3d_set_pdl(0)
        off = _solve(gpu, f0, f1, params)
        if L.flow3d_set_pdl(1) != 1:
            pytest.skip("the driver refused launches with the programmatic-serialization attribute")
        on = [_solve(gpu, f0, f1, params) for _ in range(3)]  # repeated: a race would not be deterministic
    finally:
        L.flow3d_set_pdl(-1)
    for got in on:
        for a, b in zip(got, off):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_chain_on_matches_oracle(gpu, oracle):
    L = gpu.load()
    shape = (20, 24, 30)
    params = dict(outer_iterations_count=5, inner_iterations_count=5, warp_levels_count=40)
    f0 = smooth_volume(shape, 9)
    f1 = np.ascontiguousarray(np.roll(f0, (-1, 1, 2), axis=(0, 1, 2)))
    try:
        if L.flow3d_set_pdl(1) != 1:
            pytest.skip("programmatic dependent launch unavailable")
        got = _solve(gpu, f0, f1, params)
    finally:
        L.flow3d_set_pdl(-1)
    ref = oracle.compute_flow(f0, f1, params)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
