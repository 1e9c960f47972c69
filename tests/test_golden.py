"""Parity against the reference's own CUDA build through the committed fixtures (tests/golden/README.md):
the shipped 128^3 pair and the shipped 584x388x5 slab, default parameters.  Bar: bit-exact (sha256 of the
full-resolution flows + exact equality on the committed subsample); the north-star tolerance
(max |d flow| <= 1e-3 voxel, mean endpoint-error difference <= 1e-4 voxel) is checked as well."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

META = json.load(open(os.path.join(GOLDEN, "reference_flows.json")))


def _check(name, flows, sub_file, sub_slices):
    ref = np.load(os.path.join(GOLDEN, sub_file))
    d = [np.abs(f[sub_slices] - ref[c]) for f, c in zip(flows, "uvw")]
    epe = np.sqrt(sum((f[sub_slices] - ref[c]) ** 2 for f, c in zip(flows, "uvw")))
    # north-star gate
    assert max(float(x.max()) for x in d) <= 1e-3
    assert float(epe.mean()) <= 1e-4
    # what we actually achieve: identical bits
    for f, c in zip(flows, "uvw"):
        assert np.array_equal(f[sub_slices], ref[c]), "flow_%s differs from the reference build" % c
        assert hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest() == META[name]["full_sha256"][c]


def _gpu_solve(gpu, f0, f1):
    d, h, w = f0.shape
    of = gpu.OpticalFlowE()
    of.silent = True
    assert of.Initialize(gpu.DataSize4(w, h, d))
    out = [np.zeros_like(f0) for _ in range(3)]
    of.ComputeFlow(f0, f1, out[0], out[1], out[2], dict(gpu.DEFAULTS))
    of.Destroy()
    return out


@pytest.mark.gpu
def test_gpu_pair128_matches_reference_build(gpu, pair_128):
    _check("pair128", _gpu_solve(gpu, *pair_128), "pair128_ref_flow_sub4.npz", (slice(None, None, 4),) * 3)


@pytest.mark.gpu
def test_gpu_slab_matches_reference_build(gpu, pair_slab):
    out = _gpu_solve(gpu, *pair_slab)
    _check("slab", out, "slab_ref_guarded_flow_sub4.npz", (slice(None), slice(None, None, 4), slice(None, None, 4)))


@pytest.mark.slow
def test_oracle_slab_matches_reference_build(oracle, pair_slab):
    """pins the CPU oracle against the reference's CUDA build on a full default-parameter solve"""
    out = oracle.compute_flow(*pair_slab)
    _check("slab", out, "slab_ref_guarded_flow_sub4.npz", (slice(None), slice(None, None, 4), slice(None, None, 4)))


@pytest.mark.slow
def test_oracle_pair128_matches_reference_build(oracle, pair_128):
    out = oracle.compute_flow(*pair_128)
    _check("pair128", out, "pair128_ref_flow_sub4.npz", (slice(None, None, 4),) * 3)
