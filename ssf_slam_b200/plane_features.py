"""Plane-feature extraction of the LOAM back end's first node on the GPU: what
``src/frameFeature.cpp:35-127`` computes from the ``velodyne_points`` cloud before publishing ``/plane_frame_cloud1``.
Kept next to the front end so a co-located back end can skip a host hop; the node itself is untouched."""
import torch

from . import _native as nat

# the node's parameter sets (src/frameFeature.cpp:141-151): plane_min, plane_span, rowIndexStart, rowIndexEnd
PARAMS = {16: (0.05, 3, 0, 0), 64: (0.005, 25, 5, 5)}


@torch.no_grad()
def plane_features(points, n_scan_row=16, plane_min=None, plane_span=None, row_start=None, row_end=None):
    """points f32 [B,N,3] CUDA -> (planes f32 [B,N,4] = (x,y,z,intensity) padded with zeros, count i32 [B]).
    Row b holds count[b] plane points in the node's output order (scan line ascending, position in line ascending)."""
    nat.require_device()
    d = PARAMS[n_scan_row]
    plane_min = d[0] if plane_min is None else plane_min
    plane_span = d[1] if plane_span is None else plane_span
    row_start = d[2] if row_start is None else row_start
    row_end = d[3] if row_end is None else row_end
    pts = points.contiguous().float()
    B, N, _ = pts.shape
    out = torch.zeros(B, N, 4, dtype=torch.float32, device=pts.device)
    cnt = torch.empty(B, dtype=torch.int32, device=pts.device)
    ws = torch.empty(int(nat.lib().ssf_plane_features_workspace_bytes(B, N)), dtype=torch.uint8, device=pts.device)
    nat.check(nat.lib().ssf_plane_features(nat.ptr(pts), B, N, int(n_scan_row), int(row_start), int(row_end), float(plane_min),
                                           int(plane_span), nat.ptr(ws), nat.ptr(out), nat.ptr(cnt), nat.stream()))
    return out, cnt
