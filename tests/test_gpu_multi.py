"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): tests/multigpu_worker.py under torchrun."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("voices", [256, 201, 70])
def test_master_bus_across_ranks(gpu, voices):
    """voices shard by rank; the buses are all-gathered (NCCL, side stream) and tree-summed in rank order: every rank bit-equal to the oracle"""
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, FW_TEST_VOICES=str(voices))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29700 + voices % 100), str(ROOT / "tests" / "multigpu_worker.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multigpu parity OK" in r.stdout
