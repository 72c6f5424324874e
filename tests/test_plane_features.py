"""f-3: plane-feature extraction after src/frameFeature.cpp:45-127.
Pinned: tests/golden/plane_features_ref.npz holds what the reference's OWN node publishes -- src/frameFeature.cpp compiled
unmodified into oracle/_ref/libframe_feature_ref.so (oracle/ref_build/, oracle/gen_golden_plane_features.py).
CPU: the oracle restatements (C, Python) reproduce the golden (and the compiled node itself when oracle/_ref is present).
GPU: the kernel equals the golden and the C oracle bit for bit."""
import ctypes
import os

import numpy as np
import pytest

from oracle import plane_features as opf  # checker only


def lidar_cloud(seed, n, n_rows):
    """Points on scan lines with elevations well inside a line's angular bin (scan-line decisions are then independent
    of the last ulp of atan), mixed ranges, a few degenerate points (origin, vertical, duplicates)."""
    rng = np.random.default_rng(seed)
    if n_rows == 16:
        ang = (rng.integers(-1, 17, n) * 2 - 15) + rng.uniform(-0.8, 0.8, n)
    else:
        k = rng.integers(0, 70, n)
        ang = np.where(k < 33, 2 - (k + rng.uniform(-0.4, 0.4, n)) / 3.0, -8.83 - ((k - 32) + rng.uniform(-0.4, 0.4, n)) / 2.0)
    az = np.sort(rng.uniform(-np.pi, np.pi, n))
    rad = 5 + 40 * rng.random(n) ** 2 + 0.02 * rng.standard_normal(n)
    wall = rng.random(n) < 0.5
    rad = np.where(wall, 12.0 / np.maximum(np.abs(np.cos(az)), 0.2), rad)    # a flat wall -> low curvature runs
    pts = np.stack([rad * np.cos(az), rad * np.sin(az), rad * np.tan(np.deg2rad(ang))], 1).astype(np.float32)
    pts[::97] = pts[1::97][: len(pts[::97])]      # duplicates
    pts[5] = 0                                    # origin: angle NaN -> dropped
    pts[6, :2] = 0                                # vertical: +-90 degrees -> out of range
    return pts


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "plane_features_ref.npz"))
    return [(g["pts%d" % k], int(g["rows%d" % k]), g["ref%d" % k]) for k in range(int(g["n"]))]


def test_oracle_matches_compiled_reference_node(golden_dir, oracle_c):
    """The C oracle (every cloud) and the Python oracle (the 3000-point clouds) against the outputs of the compiled reference
    node; when the node's library is present (build container; it also travels to the GPU box) it is called live as well."""
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libframe_feature_ref.so")
    lib = ctypes.CDLL(ref) if os.path.exists(ref) else None
    for pts, n_rows, want in _golden(golden_dir):
        assert np.array_equal(opf.plane_features_c(pts, n_rows), want)
        if pts.shape[0] <= 3000:
            assert np.array_equal(opf.plane_features_py(pts, n_rows), want)
        if lib is not None:
            out = np.zeros((pts.shape[0], 4), np.float32)
            cnt = ctypes.c_int(0)
            P = np.ascontiguousarray(pts, np.float32)
            assert lib.ssf_ref_plane_features(ctypes.c_void_p(P.ctypes.data), P.shape[0], n_rows, ctypes.c_void_p(out.ctypes.data), ctypes.byref(cnt)) == 0
            assert np.array_equal(out[:cnt.value], want)


@pytest.mark.gpu
def test_plane_features_gpu_equals_compiled_reference_node(golden_dir):
    import torch
    from ssf_slam_b200.plane_features import plane_features
    for pts, n_rows, want in _golden(golden_dir):
        out, cnt = plane_features(torch.from_numpy(pts[None]).cuda(), n_rows)
        out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
        assert cnt[0] == want.shape[0] and np.array_equal(out[0, :cnt[0]], want)


@pytest.mark.parametrize("n_rows,n", [(16, 700), (64, 1500)])
def test_oracle_restatements_agree(oracle_c, n_rows, n):
    pts = lidar_cloud(n_rows, n, n_rows)
    a = opf.plane_features_c(pts, n_rows)
    b = opf.plane_features_py(pts, n_rows)
    assert a.shape == b.shape and a.shape[0] > 20
    assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("n_rows,B,n", [(16, 3, 8192), (64, 2, 8192), (16, 1, 1000), (64, 1, 16384), (16, 2, 37)])
def test_plane_features_gpu_equals_oracle(oracle_c, n_rows, B, n):
    import torch
    from ssf_slam_b200.plane_features import plane_features
    clouds = np.stack([lidar_cloud(100 * n_rows + b, n, n_rows) for b in range(B)])
    out, cnt = plane_features(torch.from_numpy(clouds).cuda(), n_rows)
    out, cnt = out.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        want = opf.plane_features_c(clouds[b], n_rows)
        assert cnt[b] == want.shape[0]
        assert np.array_equal(out[b, :cnt[b]], want)
        assert not out[b, cnt[b]:].any()
