"""GPU test that needs >= 2 devices (skipped on a single-GPU box): launches tests/mgpu_allgather.py under
torchrun with 2 ranks."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_column_shard_and_nccl_allgather_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tests", "mgpu_allgather.py")],
                       capture_output=True, text=True, timeout=600)
    assert "ALLGATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
