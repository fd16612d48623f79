"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): tests/multigpu_worker.py under torchrun."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("voices,exchange,handover", [(256, "nccl", "event"), (201, "nccl", "event"), (201, "p2p", "event"), (256, "p2p", "signal")])
def test_master_bus_across_ranks(gpu, voices, exchange, handover):
    """exchange: NCCL all-gather (default) or the peer-memory push (FW_EXCHANGE=p2p) with either stream hand-over."""
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, FW_TEST_VOICES=str(voices), FW_EXCHANGE=exchange, FW_P2P_HANDOVER=handover)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29700 + voices % 100 + (7 if exchange == "p2p" else 0) + (13 if handover == "signal" else 0)), str(ROOT / "tests" / "multigpu_worker.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multigpu parity OK" in r.stdout
