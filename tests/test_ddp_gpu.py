"""GPU (needs 2 devices; skipped on a 1-GPU box): DistributedDataParallel training steps of the mirror CoordNet over the
drop-in ops with NCCL gradient all-reduce -- see tests/ddp_driver.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ddp_training_steps(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "ddp_driver.py")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, NCCL_DEBUG="WARN", NCCL_DEBUG_FILE="/dev/stderr"))
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["replicas_identical"] and d["finite"] and d["params_with_grad"] > 50 and d["world"] == 2
