import hashlib
import lzma
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def lib():
    import cuda_flow3d_b200 as pkg
    return pkg.load()


@pytest.fixture(scope="session")
def gpu():
    """the package, with a device guaranteed (fails loudly otherwise: there is no CPU fallback)"""
    import cuda_flow3d_b200 as pkg
    pkg.require_device()
    return pkg


# ---- the reference's shipped data pairs (data/*.raw, md5 in SURVEY.md section 2), xz-compressed -----
_MD5 = {
    "frame_0_128": "48c53ac7c5f7362c2392b8335c170360",
    "frame_1_128": "12ff334938378a21ca83dd17cf6c51e8",
    "rub1": "d4e924e3851be74a9636820be41ae92e",
    "rub2": "a1fc943aa4df16e4b1294c404d77b0b8",
}


def _xz(name):
    with lzma.open(os.path.join(GOLDEN, "data", name), "rb") as f:
        return f.read()


def load_pair_128():
    """(frame_0, frame_1) as float32 (128,128,128), exactly what ReadRAWFromFileU8 yields"""
    out = []
    for i in (0, 1):
        raw = _xz("frame_%d_128-128-128.raw.xz" % i)
        assert hashlib.md5(raw).hexdigest() == _MD5["frame_%d_128" % i]
        out.append(np.frombuffer(raw, np.uint8).astype(np.float32).reshape(128, 128, 128))
    return out


def load_pair_slab():
    """(rub1, rub2) as float32 (5,388,584); the shipped files are one slice repeated 5 times"""
    out = []
    for n in ("rub1", "rub2"):
        sl = np.frombuffer(_xz("%s-584-388-slice.raw.xz" % n), np.uint8).reshape(388, 584)
        vol = np.ascontiguousarray(np.broadcast_to(sl, (5, 388, 584)))
        assert hashlib.md5(vol.tobytes()).hexdigest() == _MD5[n]
        out.append(vol.astype(np.float32))
    return out


@pytest.fixture(scope="session")
def pair_128():
    return load_pair_128()


@pytest.fixture(scope="session")
def pair_slab():
    return load_pair_slab()


def random_fields(shape, seed, n, scale=1.0):
    rng = np.random.default_rng(seed)
    return [np.ascontiguousarray((rng.standard_normal(shape) * scale).astype(np.float32)) for _ in range(n)]


def smooth_volume(shape, seed):
    """band-limited positive test image"""
    rng = np.random.default_rng(seed)
    d, h, w = shape
    z, y, x = np.meshgrid(np.arange(d), np.arange(h), np.arange(w), indexing="ij")
    a = np.full(shape, 120.0)
    for _ in range(6):
        f = rng.uniform(0.02, 0.2, 3)
        p = rng.uniform(0, 6.28, 3)
        a += 18.0 * np.sin(f[0] * x + p[0]) * np.sin(f[1] * y + p[1]) * np.sin(f[2] * z + p[2])
    return np.ascontiguousarray(a.astype(np.float32))
