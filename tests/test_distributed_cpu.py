"""CPU, world_size 2 and 3 on gloo: the host-side sharding logic (slab bounds, halo, ownership, pair ranges) with
the oracle standing in for the device kernel and gloo's all_reduce standing in for the NCCL one."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_logic_gloo(world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    port = 29500 + world + (os.getpid() % 200)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "_dist_worker.py"), "cpu"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST-CPU-OK" in r.stdout


def test_slab_helpers(pkg):
    import numpy as np
    rng = np.random.default_rng(0)
    pts = np.zeros((10000, 4), np.float32)
    pts[:, :3] = rng.uniform(-50, 50, (10000, 3))
    for world in (1, 2, 4, 8):
        b = pkg.shard.slab_bounds(pts[:, 0], world)
        assert len(b) == world + 1 and b[0] == -np.inf and b[-1] == np.inf and (np.diff(b[1:-1]) > 0).all()
        own = pkg.shard.owner_of(pts[:, 0], b)
        counts = np.bincount(own, minlength=world)
        assert counts.sum() == len(pts) and counts.min() > 0.8 * len(pts) / world
        total = 0
        for r in range(world):
            sub, lo, hi, idx = pkg.shard.slab_of_rank(pts, r, world, 0, 2.0, b)
            inside = (pts[:, 0] >= lo) & (pts[:, 0] < hi)
            assert set(np.nonzero(inside)[0]) <= set(idx)          # the slab itself is complete
            near = (pts[:, 0] >= lo - 2.0) & (pts[:, 0] <= hi + 2.0)
            assert set(np.nonzero(near)[0]) <= set(idx)            # and so is the 2 m halo
            total += inside.sum()
        assert total == len(pts)
