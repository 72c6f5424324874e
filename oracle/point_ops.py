"""oracle/point_ops.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch / numpy) restatement of the point operators that SSF-SLAM's scene-flow
network calls through ``lib.pointnet2_utils`` and ``torch_scatter``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product path (``ssf_slam_b200``) never does.

The pointnet2 CUDA extension and torch_scatter are NOT in the reference tree
(``.gitignore:74``, ``README.md:22-27``; torch_scatter is an unpinned pip wheel), so each
function follows the pure-PyTorch twin the reference ships plus the arithmetic /
tie-breaking spec written down in SURVEY.md Appendix C:

* squared distance ``d = ((dx*dx) + (dy*dy)) + (dz*dz)`` in IEEE fp32, every op rounded
  (torch element-wise ops never contract to FMA) -- ``ASF/utils/utils.py:85,106``
* FPS  -- ``ASF/utils/utils.py:68-89`` with first index 0 (upstream extension) instead of
  ``torch.randint`` (``:80``); running min initialised to 1e10; argmax ties -> lowest index
* kNN  -- ``ASF/utils/utils.py:92-108``: k smallest by (d, index), ascending,
  returned distance is ``sqrt(d)``; argument order is the extension's (k, query, reference)
* ball query -- ``ASF/SetCover.py:39-63`` (the correct N-sentinel variant): ``d <= r*r``,
  ascending index, first ``nsample``, pad with first hit, ``cnt`` = hits in range
* grouping / gather -- as called at ``ASF/utils/utils.py:228-233``
* scatter_softmax / scatter_sum -- as called at ``ASF/utils/soflow.py:474,481`` (dim=1)

PINNED on tie-free data: oracle/gen_golden_point_twins.py executes the reference's own in-tree
pure-torch twins of the extension unmodified -- ``farthest_point_sample`` / ``knn_point``
(``ASF/utils/utils.py:68-108``) and ``query_ball_point`` (``ASF/SetCover.py:39-63``) -- and asserts
that both restatements here (torch and plain C) return the same indices; the vectors are committed
as tests/golden/point_twins.npz.  What stays a WRITTEN specification (the extension itself is absent
from the reference, which has no tests for it) is the behaviour on exact ties -- lowest index first --
and the no-FMA distance arithmetic; tests/golden/point_ops.npz pins those between the two restatements.
"""
import numpy as np
import torch


def _sqdist(q, r):
    """q [B,Nq,1,3] - r [B,1,Nr,3] -> [B,Nq,Nr] with the spec's summation order."""
    dx = q[..., 0] - r[..., 0]
    dy = q[..., 1] - r[..., 1]
    dz = q[..., 2] - r[..., 2]
    return (dx * dx + dy * dy) + dz * dz


# When the C restatement is built, FPS / kNN route through it (bit-identical to the torch
# versions below -- tests/test_oracle.py checks that -- and ~100x faster at N=8192).
USE_C = True


def _c_ok():
    if not USE_C:
        return False
    try:
        _clib()
        return True
    except (OSError, FileNotFoundError):
        return False


def furthest_point_sample(xyz, npoint):
    """xyz f32 [B,N,3] -> idx i32 [B,npoint]."""
    if _c_ok():
        return torch.from_numpy(c_fps(xyz.float().contiguous().numpy(), npoint))
    return furthest_point_sample_torch(xyz, npoint)


def furthest_point_sample_torch(xyz, npoint):
    xyz = xyz.float()
    B, N, _ = xyz.shape
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    mind = torch.full((B, N), 1e10, dtype=torch.float32)
    last = torch.zeros(B, dtype=torch.long)
    ar = torch.arange(B)
    for j in range(npoint):
        idx[:, j] = last.int()
        if j == npoint - 1:
            break
        c = xyz[ar, last].unsqueeze(1)  # [B,1,3]
        dx = xyz[..., 0] - c[..., 0]
        dy = xyz[..., 1] - c[..., 1]
        dz = xyz[..., 2] - c[..., 2]
        d = (dx * dx + dy * dy) + dz * dz
        mind = torch.minimum(mind, d)
        # lowest index among the maxima: first position where mind == max
        mx = mind.max(dim=1, keepdim=True)[0]
        last = (mind == mx).int().argmax(dim=1)
    return idx


def _knn_chunk(k, q, r):
    d = _sqdist(q.unsqueeze(2), r.unsqueeze(1))  # [B,nq,Nr]
    nr = d.shape[-1]
    # (d, index) lexicographic order via one unique int64 key: non-negative fp32 bit
    # patterns are monotone as integers.
    bits = d.contiguous().view(torch.int32).to(torch.int64)
    key = (bits << 32) | torch.arange(nr, dtype=torch.int64).view(1, 1, nr)
    kk, _ = torch.topk(key, k, dim=-1, largest=False, sorted=True)
    idx = (kk & 0xFFFFFFFF).to(torch.int32)
    dist = (kk >> 32).to(torch.int32).view(torch.float32)
    # torch.sqrt on CPU (SLEEF) is not correctly rounded; numpy's is (== CUDA sqrtf / sqrt.rn.f32)
    return torch.from_numpy(np.sqrt(dist.numpy())), idx


def knn(k, query, ref, chunk=1024):
    """query f32 [B,Nq,3], ref f32 [B,Nr,3] -> (dist f32 [B,Nq,k], idx i32 [B,Nq,k])."""
    if _c_ok() and k <= 64:
        d, i = c_knn(k, query.float().contiguous().numpy(), ref.float().contiguous().numpy())
        return torch.from_numpy(d), torch.from_numpy(i)
    return knn_torch(k, query, ref, chunk)


def knn_torch(k, query, ref, chunk=1024):
    query = query.float()
    ref = ref.float()
    dists, idxs = [], []
    for s in range(0, query.shape[1], chunk):
        d, i = _knn_chunk(k, query[:, s:s + chunk], ref)
        dists.append(d)
        idxs.append(i)
    return torch.cat(dists, 1), torch.cat(idxs, 1)


def three_nn(query, ref):
    return knn(3, query, ref)


def ball_query(radius, nsample, xyz, new_xyz, chunk=1024):
    """xyz f32 [B,N,3], new_xyz f32 [B,S,3] -> (idx i32 [B,S,nsample], cnt i32 [B,S])."""
    xyz = xyz.float()
    new_xyz = new_xyz.float()
    B, N, _ = xyz.shape
    r2 = torch.tensor(radius, dtype=torch.float32) * torch.tensor(radius, dtype=torch.float32)
    outs, cnts = [], []
    for s in range(0, new_xyz.shape[1], chunk):
        d = _sqdist(new_xyz[:, s:s + chunk].unsqueeze(2), xyz.unsqueeze(1))
        inr = d <= r2
        cnt = inr.sum(-1)
        cand = torch.where(inr, torch.arange(N).view(1, 1, N), torch.tensor(N))
        cand = cand.sort(dim=-1)[0][..., :nsample]
        if cand.shape[-1] < nsample:  # N < nsample
            pad = torch.full(cand.shape[:-1] + (nsample - cand.shape[-1],), N, dtype=cand.dtype)
            cand = torch.cat([cand, pad], -1)
        first = cand[..., :1]
        first = torch.where(first == N, torch.zeros_like(first), first)
        cand = torch.where(cand == N, first.expand_as(cand), cand)
        outs.append(cand.to(torch.int32))
        cnts.append(cnt.to(torch.int32))
    return torch.cat(outs, 1), torch.cat(cnts, 1)


def gather_operation(features, idx):
    """features f32 [B,C,N], idx i32 [B,M] -> [B,C,M]."""
    B, C, _ = features.shape
    return torch.gather(features, 2, idx.long().unsqueeze(1).expand(B, C, idx.shape[1]))


def grouping_operation(features, idx):
    """features f32 [B,C,N], idx i32 [B,M,S] -> [B,C,M,S]."""
    B, C, _ = features.shape
    _, M, S = idx.shape
    flat = idx.long().reshape(B, 1, M * S).expand(B, C, M * S)
    return torch.gather(features, 2, flat).reshape(B, C, M, S)


def three_interpolate(features, idx, weight):
    """features [B,C,M], idx i32 [B,N,3], weight [B,N,3] -> [B,C,N] (sum over the 3 slots in order)."""
    g = grouping_operation(features, idx)  # [B,C,N,3]
    w = weight.unsqueeze(1)
    return (g[..., 0] * w[..., 0] + g[..., 1] * w[..., 1]) + g[..., 2] * w[..., 2]


def scatter_sum(src, index, dim=1, dim_size=None):
    """src [B,L,C], index i64 [B,L] -> [B, max(index)+1, C]; rows added in ascending l."""
    assert dim == 1 and src.dim() == 3
    B, L, C = src.shape
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = torch.zeros(B, n, C, dtype=src.dtype)
    out.scatter_add_(1, index.long().unsqueeze(-1).expand(B, L, C), src)
    return out


def scatter_softmax(src, index, dim=1):
    """Segmented softmax over rows sharing an index (torch_scatter semantics)."""
    assert dim == 1 and src.dim() == 3
    B, L, C = src.shape
    n = int(index.max()) + 1
    ix = index.long().unsqueeze(-1).expand(B, L, C)
    mx = torch.full((B, n, C), float("-inf"), dtype=src.dtype)
    mx.scatter_reduce_(1, ix, src, reduce="amax", include_self=True)
    e = torch.exp(src - mx.gather(1, ix))
    den = torch.zeros(B, n, C, dtype=src.dtype).scatter_add_(1, ix, e)
    return e / den.gather(1, ix)


# ---- thin ctypes front to the C restatement (oracle/c/ssf_oracle.c), used for big sizes ----
_C = None


def _clib():
    global _C
    if _C is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c", "libssf_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " -- run `make -C oracle/c` (or __graft_entry__.build())")
        _C = ctypes.CDLL(path)
    return _C


def _p(a):
    import ctypes
    return ctypes.c_void_p(a.ctypes.data)


def c_fps(xyz, npoint):
    x = np.ascontiguousarray(xyz, np.float32)
    B, N, _ = x.shape
    out = np.empty((B, npoint), np.int32)
    _clib().ssf_oracle_fps(_p(x), B, N, npoint, _p(out))
    return out


def c_knn(k, query, ref):
    import ctypes
    q = np.ascontiguousarray(query, np.float32)
    r = np.ascontiguousarray(ref, np.float32)
    B, Nq, _ = q.shape
    Nr = r.shape[1]
    assert k <= Nr and k <= 64
    dist = np.empty((B, Nq, k), np.float32)
    idx = np.empty((B, Nq, k), np.int32)
    _clib().ssf_oracle_knn(ctypes.c_int(k), _p(q), _p(r), B, Nq, Nr, _p(dist), _p(idx))
    return dist, idx


def c_ball_query(radius, nsample, xyz, new_xyz):
    import ctypes
    x = np.ascontiguousarray(xyz, np.float32)
    c = np.ascontiguousarray(new_xyz, np.float32)
    B, N, _ = x.shape
    S = c.shape[1]
    idx = np.empty((B, S, nsample), np.int32)
    cnt = np.empty((B, S), np.int32)
    _clib().ssf_oracle_ball_query(ctypes.c_float(radius), nsample, _p(x), _p(c), B, N, S, _p(idx), _p(cnt))
    return idx, cnt


def c_group(feat, idx):
    f = np.ascontiguousarray(feat, np.float32)
    i = np.ascontiguousarray(idx, np.int32)
    B, C, N = f.shape
    _, M, S = i.shape
    out = np.empty((B, C, M, S), np.float32)
    _clib().ssf_oracle_group(_p(f), _p(i), B, C, N, M, S, _p(out))
    return out
