"""CPU-side checks of the multi-GPU boundary (include/flow3d_mgpu_c.h): the library loads next to the
single-GPU one, exports every declared symbol, and its z partition is a partition."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "flow3d_mgpu_c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flow3d_[a-z0-9_]+)\s*\(", text)))


def test_mgpu_library_exports_every_declared_symbol():
    import cuda_flow3d_b200.mgpu as m
    lib = m.load()
    declared = _declared()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert sorted(m.SIGNATURES) == declared, "python binding and header disagree"


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("d", [8, 31, 70, 95, 128, 512, 1000, 1024, 2048])
def test_own_range_is_a_balanced_partition(d, world):
    import cuda_flow3d_b200.mgpu as m
    r = [m.own_range(d, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == d
    for (a0, b0), (a1, b1) in zip(r[:-1], r[1:]):
        assert b0 == a1 and b0 > a0
    assert r[-1][1] > r[-1][0]
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 4  # cost balancing moves at most a few planes to the edge ranks
    if world > 2 and d >= 8 * world:
        assert sizes[0] >= sizes[1] and sizes[-1] >= sizes[-2]  # edge ranks own at least as many planes


def test_input_planes_cover_own_range_plus_ghost_and_blur():
    import cuda_flow3d_b200.mgpu as m
    for world in (2, 8):
        for rank in range(world):
            a, b = m.own_range(1024, rank, world)
            lo, hi = m.input_planes(1024, rank, world, 2.0, 32)
            assert lo == max(0, a - 38) and hi == min(1024, b + 38)


@pytest.mark.parametrize("shape,world,levels", [
    ((1024, 1024, 1024), 2, 40), ((1024, 1024, 1024), 8, 40), ((1024, 1024, 1024), 4, 40),
    ((2048, 2048, 2048), 8, 60), ((512, 512, 512), 2, 40), ((512, 512, 512), 8, 40),
    ((584, 388, 5), 2, 40), ((128, 128, 128), 3, 40), ((487, 301, 233), 5, 25),
    ((256, 256, 256), 2, 40), ((256, 256, 256), 4, 40), ((256, 256, 256), 8, 40),  # bench.py's N>1 parity pair
])
def test_plan_check_accepts_the_bench_geometries(shape, world, levels):
    """the partition arithmetic of the sharded solve (prolongation sources inside the previous level's valid
    planes, matching ghost sizes, all-gather pieces inside a rank's own frame slab) on every level and rank of
    the configurations bench.py and the GPU tests run (host-only dry run: no device)"""
    import cuda_flow3d_b200.mgpu as m
    ok, lv, rk = m.plan_check(*shape, world, {"warp_levels_count": levels})
    assert ok, "level %d rank %d" % (lv, rk)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_plan_check_with_test_thresholds(world):
    """the lowered thresholds the GPU tests use to shard small volumes"""
    import cuda_flow3d_b200.mgpu as m
    for shape in ((64, 48, 96), (40, 36, 120), (33, 47, 150)):
        ok, lv, rk = m.plan_check(*shape, world, {"warp_levels_count": 12}, min_planes=2, min_voxels=1)
        assert ok, "%s level %d rank %d" % (shape, lv, rk)
    # tests/test_mgpu_gpu.py: (W, H, D) = (72, 40, 48 * world), 14 levels, thresholds 8 / 1, 16 ghost planes
    ok, lv, rk = m.plan_check(72, 40, 48 * world, world, {"warp_levels_count": 14}, min_planes=8, min_voxels=1,
                              frame_ghost=16)
    assert ok, "level %d rank %d" % (lv, rk)


def test_plan_check_reports_a_frame_ghost_too_small_for_a_gathered_level():
    import cuda_flow3d_b200.mgpu as m
    # ghost of 0 planes: a coarse level's source interval (1/0.95^k planes long) reaches past the rank's slab
    ok, lv, rk = m.plan_check(256, 256, 256, 4, {"warp_levels_count": 40}, frame_ghost=0)
    assert not ok and lv >= 0 and 0 <= rk < 4


def test_frame_ghost_grows_with_the_coarsest_source_interval():
    import cuda_flow3d_b200.mgpu as m
    assert m.frame_ghost(1024, 1024, 1024, 8) == 32                             # default pyramid
    assert m.frame_ghost(2048, 2048, 2048, 8, {"warp_levels_count": 60}) == 32  # BASELINE configs[4]
    g = m.frame_ghost(512, 512, 512, 4, {"warp_scale_factor": 0.9})             # coarsest interval: 61 planes
    assert g == 64 and m.plan_check(512, 512, 512, 4, {"warp_scale_factor": 0.9}, frame_ghost=g)[0]
    assert not m.plan_check(512, 512, 512, 4, {"warp_scale_factor": 0.9}, frame_ghost=32)[0]
    assert m.frame_ghost(300, 200, 40, 2, {"warp_scale_factor": 0.5, "warp_levels_count": 10}) <= 40  # capped at the depth
    assert m.frame_ghost(64, 64, 64, 1) == 32
    # bench.py N>1: default runs keep the 32 planes every recorded run used; the parity pair (thresholds 8 / 1) too
    for world in (2, 4, 8):
        assert m.frame_ghost(1024, 1024, 1024, world) == 32
        assert m.frame_ghost(256, 256, 256, world, None, 8, 1) == 32
    assert m.frame_ghost(2048, 2048, 2048, 8, {"warp_levels_count": 80}) == 64


def test_plan_passes_whenever_the_ghost_depth_covers_the_median():
    """documented in flow3d_mgpu_c.h: with whole-frame ghosts, inner >= median_radius // 2 + 1 is sufficient for every
    geometry (sampled); below that the plan check refuses some geometries up front instead of mid-solve"""
    import random
    import cuda_flow3d_b200.mgpu as m
    rnd = random.Random(11)
    refused_below = 0
    for _ in range(1500):
        W, H, D = rnd.randint(8, 300), rnd.randint(8, 300), rnd.randint(40, 2100)
        world = rnd.choice([2, 3, 4, 8])
        med = rnd.choice([1, 3, 5, 7])
        inner = rnd.randint(0, 7)
        P = {"warp_levels_count": 40, "warp_scale_factor": rnd.choice([0.5, 0.8, 0.95]), "inner_iterations_count": inner,
             "median_radius": med}
        ok, lv, rk = m.plan_check(W, H, D, world, P, min_planes=2, min_voxels=1, frame_ghost=D)
        if inner >= med // 2 + 1:
            assert ok, (W, H, D, world, P, lv, rk)
        else:
            refused_below += int(not ok)
    assert refused_below > 0
