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
