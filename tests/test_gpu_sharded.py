"""GPU: slab-sharded registration (C5 shape, scaled down).

test_slabs_on_one_gpu runs on ANY box: the ranks are emulated one after the other on one device (never as kernels
that wait for one another) — every slab's map (slab + halo) is built in turn, the device kernel of the sharded loop
(ownership pre-pass, owned-point lists, search in the slab, accumulation; MODE_ITER_NOSOLVE) produces that slab's
partial accumulators through icp4r_accumulate_slab, the host adds them up like the all-reduce would, and the sum must
equal the unsharded accumulators of the device and of the oracle. It then iterates that loop to a pose and compares
it with icp4r_register_map on the whole map.

test_register_sharded_two_ranks needs >= 2 devices (NCCL cannot place two ranks on one device): the real
icp4r_register_sharded with the NCCL all-reduce and with the in-kernel peer-memory exchange."""
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


def test_slabs_on_one_gpu(pkg, O, handle):
    import numpy as np
    gate = 2.0
    scan, mp, _ = pkg.synth.scan_to_map(77, 1500, 60000, extent=60.0, scan_radius=50.0)
    world = 3
    bounds = pkg.shard.slab_bounds(mp[:, 0], world)
    slabs = [pkg.shard.slab_of_rank(mp, r, world, axis=0, halo=gate, bounds=bounds) for r in range(world)]
    for kind, k in ((pkg.P2PLANE_KNN, 5), (pkg.P2P_SVD, 1), (pkg.P2LINE, 2)):
        o = pkg.default_opts(residual=kind, k=k, max_iterations=6, max_corr_dist=gate)
        oo = O.default_opts(residual=kind, k=k, max_iterations=6, max_corr_dist=gate)
        nv = 17 if kind == pkg.P2P_SVD else 29
        T = pkg.synth.se3(0.01, 0, 0, (0.2, -0.1, 0.0))
        total = np.zeros(pkg.api.ACC_LEN)
        owned = 0
        for mine, lo, hi, _gidx in slabs:
            handle.map_build(mine)
            part = handle.accumulate_slab(scan, o, T, 0, lo, hi)
            owned += int(part[0] if kind == pkg.P2P_SVD else part[28])
            total += part
        handle.map_build(mp)
        whole = handle.accumulate_slab(scan, o, T, -1)
        want, _idx, used = O.accumulate(scan, mp, oo, T)
        assert owned == int(whole[0] if kind == pkg.P2P_SVD else whole[28]) == used
        assert np.allclose(total[:nv], whole[:nv], rtol=1e-12, atol=1e-9), np.abs(total[:nv] - whole[:nv]).max()
        assert np.linalg.norm(total[:nv] - want[:nv]) <= 1e-9 * np.linalg.norm(want[:nv])
    # the sharded LOOP, emulated: per iteration every slab's pass, host sum, the oracle's solve on the summed accumulators
    o = pkg.default_opts(residual=pkg.P2PLANE_KNN, k=5, max_iterations=6, max_corr_dist=gate)
    oo = O.default_opts(residual=O.P2PLANE_KNN, k=5, max_iterations=6, max_corr_dist=gate)
    T = np.eye(4)
    for _ in range(6):
        total = np.zeros(pkg.api.ACC_LEN)
        for mine, lo, hi, _gidx in slabs:
            handle.map_build(mine)
            total += handle.accumulate_slab(scan, o, T, 0, lo, hi)
        T = O.gn_update(total, T)
    handle.map_build(mp)
    T1, r1, _ = handle.register_map(scan, o)
    assert np.abs(T - T1).max() < 1e-9, np.abs(T - T1).max()
