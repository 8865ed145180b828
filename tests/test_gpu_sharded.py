"""Multi-GPU: point-sharded registration (NCCL all-reduce of the packed normal equations per evaluation) equals the
single-GPU result. Needs >= 2 GPUs; skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_single_gpu():
    from rolo_b200 import capi
    n = capi.load().rg_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "sharded_check.py"), "small"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 3      # fused NVLink P2P reduction, NCCL on device buffers, gloo on host buffers
