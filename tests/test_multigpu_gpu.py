"""The sharded path on real GPUs: two NCCL ranks (one process per GPU) run the image-sharded sweep + embed + winners-only
exchange through the public API; values, ids and the concept DB must equal a single-process run bit for bit
(scripts/check_multigpu.py asserts it on every rank). Skipped when fewer than two GPUs are visible."""

import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("exchange", ["winners", "all"])
def test_two_rank_concept_db_equals_single_rank(exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "scripts" / "check_multigpu.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=850, cwd=ROOT,
                         env=dict(os.environ, SLB_EXCHANGE=exchange))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("multi-gpu parity ok") == 2
