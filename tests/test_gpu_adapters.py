"""GPU: the C++ adapters that mirror the reference's two call shapes (KD_TREE<PointType> and the PCL / fast_gicp
Registration objects) run against libicp4r_cuda and check themselves against an exhaustive search."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_adapters_selftest():
    d = os.path.join(ROOT, "icp-4dradar_b200", "adapters")
    exe = os.path.join(d, "test_adapters")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", d], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adapters ok" in r.stdout


def test_plain_c_example_runs():
    """the plain-C host program (C99, no CUDA headers) drives the scan-to-map loop through the C ABI alone"""
    d = os.path.join(ROOT, "icp-4dradar_b200", "adapters")
    exe = os.path.join(d, "example_c_abi")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", d], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("frame ")]
    assert len(lines) == 10 and "map 30000 points" in lines[-1]
    # the synthetic walls make x observable: the last frame's estimate is within 10 cm of the 3.6 m travelled
    x = float(lines[-1].split("x = ")[1].split(" m")[0])
    assert abs(x - 3.6) < 0.1, lines[-1]
