"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 through the C ABI."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402  (tests may use the oracle; the product may not)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def port():
    return pyoracle.Checker("port")


@pytest.fixture(scope="session")
def ref():
    if not pyoracle.have_reference():
        pytest.skip("oracle/_ref/libref_harness.so not built (needs /root/reference: make -C oracle ref)")
    return pyoracle.Checker("reference")


class HostBuild(pyoracle.Checker):
    """g++ build of the product's bit logic (tests/host_build/hostlib.cpp) behind the Checker API."""

    def __init__(self):
        d = os.path.join(ROOT, "tests", "host_build")
        so = os.path.join(d, "libb2p_hosttest.so")
        srcs = [os.path.join(d, "hostlib.cpp")] + [os.path.join(ROOT, "gpu_ai_b200", "csrc", f)
                                                   for f in ("bitboard.cuh", "philox.cuh", "playout_core.cuh")]
        defs = os.environ.get("B2P_HOST_DEFS", "").split()   # experiment flags (e.g. -DB2P_MASK_TARGETS): same tests
        if defs:
            so = os.path.join(d, "libb2p_hosttest_exp.so")
        if defs or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared", "-o", so,
                            srcs[0]] + defs, check=True)
        self.kind = "hostbuild"
        self.lib = C.CDLL(so)
        self.pfx = "hb_"


@pytest.fixture(scope="session")
def hostbuild():
    return HostBuild()


@pytest.fixture(scope="session")
def engine():
    import gpu_ai_b200
    return gpu_ai_b200.Engine(devices=1, seed=12345)


def unflatten(flat, counts, max_moves=64):
    out = np.zeros((len(counts), max_moves), dtype=np.uint64)
    pos = 0
    for i, c in enumerate(int(x) for x in counts):
        out[i, :c] = flat[pos:pos + c]
        pos += c
    return out


def synthetic_positions(n, seed):
    """Same generator as tools/make_golden.py (king-rich random placements)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 4), dtype=np.uint32)
    for i in range(n):
        total = int(rng.integers(2, 21))
        squares = rng.choice(32, size=total, replace=False)
        n1 = int(rng.integers(1, total))
        p1 = p2 = k = 0
        for j, s in enumerate(squares):
            if j < n1:
                p1 |= 1 << int(s)
            else:
                p2 |= 1 << int(s)
            if rng.random() < 0.5:
                k |= 1 << int(s)
        msc = int(rng.choice([0, 1, 10, 48, 49, 50, 51]))
        out[i] = (p1, p2, k, int(rng.integers(0, 2)) | (msc << 8))
    return out


def fast_synthetic(n, seed):
    """Vectorised king-rich random placements for the million-position sweeps."""
    rng = np.random.default_rng(seed)
    occ = rng.random((n, 32)) < rng.uniform(0.05, 0.6, size=(n, 1))
    side = rng.random((n, 32)) < 0.5
    king = rng.random((n, 32)) < rng.uniform(0.0, 1.0, size=(n, 1))
    w = (1 << np.arange(32, dtype=np.uint64))
    p1 = ((occ & side) * w).sum(axis=1).astype(np.uint32)
    p2 = ((occ & ~side) * w).sum(axis=1).astype(np.uint32)
    k = ((occ & king) * w).sum(axis=1).astype(np.uint32)
    meta = (rng.integers(0, 2, n) | (rng.choice([0, 3, 47, 49, 50], n) << 8)).astype(np.uint32)
    return np.stack([p1, p2, k, meta], axis=1)
