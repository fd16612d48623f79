import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    return pyoracle.load()


@pytest.fixture(scope="session")
def product():
    import firewheel_b200
    lib = firewheel_b200.load()  # raises if the CUDA library is missing: no CPU fallback
    return lib


@pytest.fixture(scope="session")
def gpu(product):
    n = product.device_count()
    if n <= 0:
        pytest.fail("GPU test selected but no CUDA device is visible: " + (product.last_device_error() or b"").decode())
    return product


def pcg32_uniform(seed, n):
    """SURVEY §8d synthetic input: x = ((pcg32(seed) >> 8) * 2^-24) * 2 - 1, as float32."""
    # vectorised PCG32 (XSH-RR) with a fixed stream
    mult, inc = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    state = np.uint64(seed) * mult + inc
    out = np.empty(n, dtype=np.float32)
    with np.errstate(over="ignore"):
        for i in range(n):
            old = state
            state = old * mult + inc
            xorshifted = np.uint32(((old >> np.uint64(18)) ^ old) >> np.uint64(27))
            rot = np.uint32(old >> np.uint64(59))
            r = (xorshifted >> rot) | (xorshifted << ((~rot + np.uint32(1)) & np.uint32(31)))
            out[i] = np.float32((int(r) >> 8) * (2.0 ** -24) * 2.0 - 1.0)
    return out


def synth(shape, seed):
    """Fast seeded uniform[-1,1) float32 with exactly-representable 24-bit mantissas (numpy PCG64)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    r = rng.integers(0, 1 << 24, size=shape, dtype=np.uint32)
    return (r.astype(np.float32) * np.float32(2.0 ** -24) * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
