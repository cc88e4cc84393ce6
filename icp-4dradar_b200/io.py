"""File formats of the reference's scan-to-scan node, so recorded sequences can be replayed through the library.

* ``radar_pointcloud_<n>.bin`` frames: a flat float32 array, five values per point in the order
  x, y, z, intensity, doppler (read_radar_data, /root/reference/src/iterative_closest_point.cpp:64-82; field order
  :373-377; file naming :303-304).
* ``icp.txt`` / ``icp_map.txt``: one pose per line, the upper 3x4 of the 4x4 transform row-major, 12 values separated
  by spaces (/root/reference/src/iterative_closest_point.cpp:778-789).
* ``output_result.csv`` rows: time, the 16 entries of T row-major, score, A, b (:191,701-706).
"""
from __future__ import annotations

import os

import numpy as np


def frame_path(folder: str, index: int) -> str:
    return os.path.join(folder, f"radar_pointcloud_{index}.bin")


def read_radar_bin(path: str) -> np.ndarray:
    """[n,5] float32 (x, y, z, intensity, doppler); a missing file yields an empty frame like the reference (:77-81)."""
    if not os.path.exists(path):
        return np.zeros((0, 5), np.float32)
    a = np.fromfile(path, dtype=np.float32)
    return a[: (a.size // 5) * 5].reshape(-1, 5)


def write_radar_bin(path: str, records: np.ndarray) -> None:
    np.ascontiguousarray(records, np.float32).reshape(-1, 5).tofile(path)


def xyzw_of(records: np.ndarray) -> np.ndarray:
    """the packed x, y, z, intensity rows the registration entry points take"""
    return np.ascontiguousarray(records[:, :4], np.float32)


def write_icp_txt(path: str, poses) -> None:
    with open(path, "w") as f:
        for T in poses:
            T = np.asarray(T, np.float64).reshape(4, 4)
            f.write(" ".join(repr(float(v)) for v in T[:3, :].reshape(-1)) + "\n")


def read_icp_txt(path: str):
    out = []
    with open(path) as f:
        for line in f:
            v = [float(x) for x in line.split()]
            if len(v) != 12:
                continue
            T = np.eye(4)
            T[:3, :] = np.asarray(v).reshape(3, 4)
            out.append(T)
    return out


def result_csv_row(time_s: float, T, score: float, A: float, b: float) -> str:
    T = np.asarray(T, np.float64).reshape(16)
    return ",".join([repr(float(time_s))] + [repr(float(v)) for v in T] + [repr(float(score)), repr(float(A)), repr(float(b))])
