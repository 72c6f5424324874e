"""Drop-in for the reference's ``lib.pointnet2_utils`` (the pointnet2 CUDA extension the reference imports at
``ASF/utils/utils.py:7``, ``ASF/utils/soflow.py:7`` but does not ship: ``.gitignore:74``, ``README.md:22-27``).

Same names, argument order, shapes and dtypes as the reference's call sites
(``ASF/utils/utils.py:226-233,291-302``; ``ASF/utils/soflow.py:30,387-406,1241-1249,1459-1470``); results are
torch CUDA tensors allocated on the current stream.  Every function launches a hand-written sm_100a kernel
through the C ABI (``include/ssf_b200.h``); CPU tensors raise -- there is no fallback.  Inference only (no
autograd), which is all the reference's ROS drivers use (``torch.no_grad()``, ``net.eval()``).

To run unmodified reference files put ``ssf_slam_b200/compat`` on ``sys.path``: it provides the packages ``lib``
and ``torch_scatter`` re-exporting this module and ``ssf_slam_b200.scatter``.
"""
import torch

from . import _native as nat
from . import functional as F_


def _f32(t):
    if t.dtype != torch.float32:
        raise nat.SsfError("expected a float32 tensor, got %s" % t.dtype)
    return t.contiguous()


def _i32(t):
    if t.dtype != torch.int32:
        raise nat.SsfError("expected an int32 index tensor, got %s (the reference passes knn_idx.int())" % t.dtype)
    return t.contiguous()


@torch.no_grad()
def furthest_point_sample(xyz, npoint):
    """xyz f32 [B,N,3] -> idx i32 [B,npoint] (start index 0, ties -> lowest index)."""
    nat.require_device()
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    nat.check(nat.lib().ssf_furthest_point_sample(nat.ptr(xyz), B, N, int(npoint), nat.ptr(out), nat.stream()))
    return out


@torch.no_grad()
def gather_operation(features, idx):
    """features f32 [B,C,N], idx i32 [B,M] -> f32 [B,C,M]."""
    nat.require_device()
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M, dtype=torch.float32, device=features.device)
    nat.check(nat.lib().ssf_gather_operation(nat.ptr(features), nat.ptr(idx), B, C, N, M, nat.ptr(out), nat.stream()))
    return out


INDEX_MAX_POINTS = 131072


@torch.no_grad()
def build_index(xyz):
    """Spatial index of the clouds xyz f32 [B,N,3] (N <= 131072) for ``knn(..., index=)`` and ``ball_query(..., index=)``:
    Hilbert-curve-sorted blocks of 32 points with bounding boxes (csrc/knn_blocks.cu).  Extension of the reference API: it lets
    several searches in the same cloud (a radius sweep, kNN + ball query) share one build; every search returns exactly what
    the un-indexed operator returns."""
    nat.require_device()
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    L = nat.lib()
    ws = torch.empty(int(L.ssf_knn_blocks_workspace_floats(B, N)), dtype=torch.float32, device=xyz.device)
    nat.check(L.ssf_knn_blocks_build(nat.ptr(xyz), B, N, nat.ptr(ws), nat.stream()))
    return ws


@torch.no_grad()
def knn(k, unknown, known, offset=None, index=None):
    """(k, query f32 [B,Nq,3], reference f32 [B,Nr,3]) -> (dist f32 [B,Nq,k] ascending, idx i32 [B,Nq,k]).
    ``offset`` (extension): query is taken as unknown + offset without materialising the sum.
    ``index`` (extension): ``build_index(known)``, to share the build between calls."""
    nat.require_device()
    unknown, known = _f32(unknown), _f32(known)
    B, Nq, _ = unknown.shape
    Nr = known.shape[1]
    dist = torch.empty(B, Nq, k, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, Nq, k, dtype=torch.int32, device=unknown.device)
    off = None if offset is None else _f32(offset)
    L = nat.lib()
    if index is not None or F_.KNN_BLOCKS_MIN_REF <= Nr <= 16384 or (16384 < Nr <= INDEX_MAX_POINTS and B * Nq > 32768):
        # block search (csrc/knn_blocks.cu): bit-identical to the brute-force scan, several times faster; above 16384
        # reference points a two-level index (radix-sorted build, super-blocks of 32 blocks)
        ws = index if index is not None else build_index(known)
        nat.check(L.ssf_knn_blocks_search(int(k), nat.ptr(unknown), nat.ptr(off), nat.ptr(ws), B, Nq, Nr, nat.ptr(dist),
                                          nat.ptr(idx), nat.stream()))
        return dist, idx
    if Nr > 16384 and B * Nq <= 32768:
        # few queries against a large cloud: one warp per two queries (one thread per query would leave the GPU empty)
        nat.check(L.ssf_knn_warp_scan(int(k), nat.ptr(unknown), nat.ptr(off), nat.ptr(known), B, Nq, Nr, nat.ptr(dist),
                                      nat.ptr(idx), nat.stream()))
        return dist, idx
    nat.check(L.ssf_knn_offset(int(k), nat.ptr(unknown), nat.ptr(off), nat.ptr(known), B, Nq, Nr, nat.ptr(dist),
                                nat.ptr(idx), nat.stream()))
    return dist, idx


@torch.no_grad()
def three_nn(unknown, known):
    return knn(3, unknown, known)


@torch.no_grad()
def ball_query(radius, nsample, xyz, new_xyz, return_count=False, index=None, use_index=None):
    """(radius, nsample, xyz f32 [B,N,3], new_xyz f32 [B,S,3]) -> idx i32 [B,S,nsample] (, cnt i32 [B,S]).
    Clouds of 2048 .. 131072 points are searched through the spatial index (``index`` = ``build_index(xyz)`` to share the build
    between calls, e.g. a radius sweep); smaller ones, nsample > 32 or ``use_index=False`` by the brute-force scan.  Same output."""
    nat.require_device()
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = torch.empty(B, S, nsample, dtype=torch.int32, device=xyz.device)
    cnt = torch.empty(B, S, dtype=torch.int32, device=xyz.device) if return_count else None   # without counts a query stops at nsample hits
    if use_index is None:
        use_index = index is not None or (2048 <= N <= INDEX_MAX_POINTS and 0 < nsample <= 32)
    if use_index:
        ws = index if index is not None else build_index(xyz)
        nat.check(nat.lib().ssf_ball_query_blocks(float(radius), int(nsample), nat.ptr(new_xyz), nat.ptr(ws), B, N, S, nat.ptr(idx),
                                                  nat.ptr(cnt), nat.stream()))
        return (idx, cnt) if return_count else idx
    nat.check(nat.lib().ssf_ball_query(float(radius), int(nsample), nat.ptr(xyz), nat.ptr(new_xyz), B, N, S, nat.ptr(idx),
                                       nat.ptr(cnt), nat.stream()))
    return (idx, cnt) if return_count else idx


@torch.no_grad()
def grouping_operation(features, idx):
    """features f32 [B,C,N], idx i32 [B,M,S] -> f32 [B,C,M,S]."""
    nat.require_device()
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    _, M, S = idx.shape
    out = torch.empty(B, C, M, S, dtype=torch.float32, device=features.device)
    nat.check(nat.lib().ssf_grouping_operation(nat.ptr(features), nat.ptr(idx), B, C, N, M, S, nat.ptr(out), nat.stream()))
    return out


@torch.no_grad()
def three_interpolate(features, idx, weight):
    """features f32 [B,C,M], idx i32 [B,N,3], weight f32 [B,N,3] -> f32 [B,C,N]."""
    nat.require_device()
    features, idx, weight = _f32(features), _i32(idx), _f32(weight)
    B, C, M = features.shape
    N = idx.shape[1]
    out = torch.empty(B, C, N, dtype=torch.float32, device=features.device)
    nat.check(nat.lib().ssf_three_interpolate(nat.ptr(features), nat.ptr(idx), nat.ptr(weight), B, C, M, N, nat.ptr(out),
                                              nat.stream()))
    return out


class QueryAndGroup(torch.nn.Module):
    """Constructible as in the reference (only instantiated in dead code, ASF/utils/soflow.py:1520-1523)."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped = grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            return grouped
        gf = grouping_operation(features, idx)
        return torch.cat([grouped, gf], dim=1) if self.use_xyz else gf


class GroupAll(torch.nn.Module):
    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped
        gf = features.unsqueeze(2)
        return torch.cat([grouped, gf], dim=1) if self.use_xyz else gf
