"""GPU, needs >= 2 devices: slab-sharded registration (C5 shape, scaled down) through icp4r_register_sharded with the
NCCL all-reduce of the 29 accumulators. Skipped on single-GPU boxes (NCCL cannot place two ranks on one device);
the host logic is covered on CPU by tests/test_distributed_cpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_register_sharded_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", os.path.join(ROOT, "tests", "_dist_worker.py"), "gpu"],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST-GPU-OK" in r.stdout and "DIST-GPU-FUSED-OK" in r.stdout, r.stdout[-2000:]
