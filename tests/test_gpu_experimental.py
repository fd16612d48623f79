"""Opt-in validation of kernels that are compiled but not yet enabled by default. Skipped unless FW_VALIDATE_EXPERIMENTAL=1.

* FW_TEMPORAL_WS=1 — the warp-specialised (compute warp + mover warp) variant of the temporal lanes kernel
  (firewheel_b200/csrc/temporal.cu). The knob is read once per process, so the temporal parity tests are re-run in a child
  process with the knob set; they compare against the oracle bit for bit, exactly as for the default kernel."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("FW_VALIDATE_EXPERIMENTAL") != "1", reason="opt-in: FW_VALIDATE_EXPERIMENTAL=1")]
ROOT = Path(__file__).resolve().parent.parent


def test_warp_specialised_temporal_kernel_is_bit_exact():
    env = dict(os.environ, FW_TEMPORAL_WS="1")
    env.pop("FW_VALIDATE_EXPERIMENTAL")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-k", "biquad or temporal or config3 or config5 or svf or golden or full_size",
                        str(ROOT / "tests")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
