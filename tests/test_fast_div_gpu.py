"""The sweep kernels' branch-free division (csrc/common.cuh: div_fast = the fast path of div.rn.f32 without
its range-check branch) must equal IEEE division bit for bit on every operand pair it accepts: checked on
the device against __fdiv_rn over every divisor mantissa and 2^32 random pairs."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def _run(lib, n, seed, mode):
    out = (C.c_uint64 * 3)()
    assert lib.flow3d_selftest_fast_div(n, seed, mode, out) == 0
    return [int(x) for x in out]


def test_fast_div_random_pairs(gpu, lib):
    bad, rejected, tested = _run(lib, 1 << 32, 12345, 0)
    assert tested == 1 << 32
    assert bad == 0
    assert rejected == 0  # the generator stays inside the accepted range


def test_fast_div_every_divisor_mantissa(gpu, lib):
    for seed in (1, 2, 3):
        bad, rejected, tested = _run(lib, 64 << 23, seed, 1)
        assert tested == 64 << 23 and bad == 0 and rejected == 0


def test_fast_sqrt_and_rcp_every_float(gpu, lib):
    """sqrt_fast / rcp_fast (the robust weights 1 / (2 sqrt(s))) against __fsqrt_rn / __frcp_rn on every float
    bit pattern: no accepted argument may differ; the accepted set must cover the normal range they are used on"""
    for mode in (2, 3):
        bad, rejected, tested = _run(lib, 1 << 32, 0, mode)
        assert tested == 1 << 32 and bad == 0
        assert rejected < (1 << 32) * 0.6  # sqrt rejects negatives, tiny values, Inf/NaN; rcp rejects few
