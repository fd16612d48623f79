"""Committed golden vectors (tests/golden/*.npz, rendered by tests/golden/make_golden.py from the CPU oracle).
CPU half: the oracle built here reproduces them bit for bit — pins the restatement against compiler / flag drift.
GPU half: the CUDA product reproduces them with no oracle in the loop (bit-exact; the reverb within 1e-5)."""
from pathlib import Path

import numpy as np
import pytest

import golden_scenarios as gs

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name):
    d = np.load(GOLDEN / f"{name}.npz")
    n = sum(1 for k in d.files if k.startswith("in"))
    inputs = [d[f"in{i}"] for i in range(n)]
    extra = {k[len("extra_"):]: d[k] for k in d.files if k.startswith("extra_")}
    want = [(d[f"out{i}"], int(d[f"mask{i}"])) for i in range(n)]
    return inputs, extra, want


def check(lib, name):
    inputs, extra, want = load(name)
    tol = gs.SCENARIOS[name][1]
    got = gs.run(lib, name, inputs, extra)
    assert len(got) == len(want)
    for i, ((y, m), (yw, mw)) in enumerate(zip(got, want)):
        if tol is None:
            assert np.array_equal(y.view(np.uint32), yw.view(np.uint32)), f"{name} call {i}: {int(np.sum(y.view(np.uint32) != yw.view(np.uint32)))} samples differ"
        else:
            err = float(np.max(np.abs(y.astype(np.float64) - yw)) / np.max(np.abs(yw)))
            assert err <= tol, (name, i, err)
        assert m == mw, (name, i, hex(m), hex(mw))


@pytest.mark.parametrize("name", sorted(gs.SCENARIOS))
def test_oracle_reproduces_golden(oracle, name):
    check(oracle, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(gs.SCENARIOS))
def test_product_reproduces_golden(gpu, name):
    check(gpu, name)
