"""CPU: the reference's frame / trajectory file formats round-trip."""
import numpy as np


def test_bin_and_icp_txt_roundtrip(pkg, tmp_path):
    rng = np.random.default_rng(0)
    rec = rng.normal(size=(137, 5)).astype(np.float32)
    p = pkg.io.frame_path(str(tmp_path), 12)
    assert p.endswith("radar_pointcloud_12.bin")
    pkg.io.write_radar_bin(p, rec)
    assert open(p, "rb").read() == rec.tobytes()                 # flat float32, 5 per point, as read_radar_data expects
    back = pkg.io.read_radar_bin(p)
    assert back.shape == (137, 5) and (back == rec).all()
    assert pkg.io.read_radar_bin(str(tmp_path / "missing.bin")).shape == (0, 5)
    assert (pkg.io.xyzw_of(rec) == rec[:, :4]).all()
    poses = [pkg.synth.se3(0.1 * i, 0.01, -0.02, (i, -i, 0.5)) for i in range(5)]
    q = str(tmp_path / "icp.txt")
    pkg.io.write_icp_txt(q, poses)
    lines = open(q).read().strip().splitlines()
    assert len(lines) == 5 and all(len(l.split()) == 12 for l in lines)   # 3x4 row-major per line
    for a, b in zip(pkg.io.read_icp_txt(q), poses):
        assert np.array_equal(a, b)
    row = pkg.io.result_csv_row(1.5, poses[1], 0.25, -2.0, 0.1).split(",")
    assert len(row) == 20 and float(row[0]) == 1.5 and float(row[17]) == 0.25
