"""SURVEY 8e on real GPUs (skipped on a single-GPU box): two ranks under torchrun run the learner with (A) the NCCL protocol and
(B) the NVLink peer-memory exchange inside the library; B == A to summation-order noise, all ranks hold identical parameters,
and both match the fp64 full-batch oracle (tools/multi_gpu_check.py).  The host-side protocol is covered without GPUs by
tests/test_host_logic.py (gloo, world size 2)."""
import os
import random
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_memory_exchange_equals_nccl_and_full_batch_oracle():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs on one node")
    port = random.randint(20000, 40000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI-GPU OK world=2" in out.stdout
