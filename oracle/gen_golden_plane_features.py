"""oracle/gen_golden_plane_features.py -- TEST INFRASTRUCTURE.  Run in the BUILD container only (needs /root/reference).

Pins the plane-feature oracle (oracle/plane_features.py, oracle/c) against the reference's OWN compiled node: builds
oracle/_ref/libframe_feature_ref.so from /root/reference/src/frameFeature.cpp (unmodified, compiled in place against the
stand-in ROS / PCL headers of oracle/ref_build/stubs; recipe oracle/ref_build/Makefile), feeds its cloudHandler clouds of three
kinds -- scan-line shaped, CARLA-shaped synthetic frames, uniform full-beam clouds (the latter two do NOT keep elevations away
from the scan-line bin edges) -- asserts that both oracle restatements reproduce what the node publishes on /plane_frame_cloud1
bit for bit, and commits inputs + reference outputs to tests/golden/plane_features_ref.npz.

Usage:  python -m oracle.gen_golden_plane_features
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import plane_features as opf  # noqa: E402
from ssf_slam_b200 import synth  # noqa: E402


def reference_lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref_build")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libframe_feature_ref.so"))
    lib.ssf_ref_plane_features.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def reference_plane_features(lib, points, n_rows):
    P = np.ascontiguousarray(points, np.float32)
    out = np.zeros((P.shape[0], 4), np.float32)
    cnt = ctypes.c_int(0)
    assert lib.ssf_ref_plane_features(P.ctypes.data, P.shape[0], n_rows, out.ctypes.data, ctypes.byref(cnt)) == 0
    return out[:cnt.value].copy()


def clean_lidar(seed, n, n_rows):
    """Exact scan lines (elevation at the bin centre), azimuth-ordered, smooth ranges (two walls and a curved facade): long
    low-curvature runs, so the greedy selection with its plane_span skip -- not just the line ends -- is exercised."""
    rng = np.random.default_rng(seed)
    if n_rows == 16:
        ang = rng.integers(0, 16, n) * 2.0 - 15.0
    else:
        k = rng.integers(0, 64, n)
        ang = np.where(k < 33, 2 - k / 3.0, -8.83 - (k - 32) / 2.0)
    az = np.sort(rng.uniform(-np.pi, np.pi, n))
    rad = np.where(np.abs(np.sin(az)) > 0.5, 9.0 / np.maximum(np.abs(np.sin(az)), 0.5), 14.0 + 2.0 * np.cos(3 * az))
    return np.stack([rad * np.cos(az), rad * np.sin(az), rad * np.tan(np.deg2rad(ang))], 1).astype(np.float32)


def main():
    from test_plane_features import lidar_cloud
    lib = reference_lib()
    out, k = {}, 0
    for n_rows in (16, 64):
        clouds = [lidar_cloud(7 + n_rows, 8192, n_rows), clean_lidar(8 + n_rows, 3000, n_rows), clean_lidar(9 + n_rows, 12000, n_rows),
                  synth.make_pair(61 + n_rows, 8192)["pos2"], synth.dense_cloud(n_rows, 12000)]
        for pts in clouds:
            want = reference_plane_features(lib, pts, n_rows)
            assert np.array_equal(opf.plane_features_c(pts, n_rows), want), "C oracle deviates from the compiled reference node"
            if pts.shape[0] <= 3000:
                assert np.array_equal(opf.plane_features_py(pts, n_rows), want), "Python oracle deviates"
            out["pts%d" % k], out["rows%d" % k], out["ref%d" % k] = pts.astype(np.float32), n_rows, want
            print("cloud %d: n_rows %d, %d points -> %d plane points: reference node == oracle" % (k, n_rows, pts.shape[0], want.shape[0]))
            k += 1
    out["n"] = k
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "plane_features_ref.npz"), **out)
    print("wrote tests/golden/plane_features_ref.npz")


if __name__ == "__main__":
    main()
