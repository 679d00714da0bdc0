"""Two-GPU test of the peer-memory exchange of config C3 (skipped on a single-GPU box): scripts/check_peer_exchange.py under
torchrun compares it with the NCCL path and with an unsharded engine."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_memory_exchange_matches_nccl_and_unsharded():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "check_peer_exchange.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, R_TOTAL="2048", ITERS="4"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "peer exchange OK" in out.stdout
