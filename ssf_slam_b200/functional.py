"""Point-major functional layer of the fused path: thin Python over the C ABI (``include/ssf_b200.h``).

Tensors are torch CUDA fp32 ``[B, N, C]`` (features) / ``[B, N, 3]`` (coordinates) / int32 ``[B, N, k]`` (indices);
weights are K-major ``[Cin, Cout]`` (see ``ssf_slam_b200.model.prepare_weights``).  Every function is one kernel
launch of ours on the current stream; nothing here computes with torch ops.
"""
import ctypes
import threading

import torch

from . import _native as nat

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
# tcgen05 realisation of the dense layers where one exists (False: SIMT fp32 kernels everywhere; used by the tests
# that compare the two realisations)
USE_TC = True


def _rows(x):
    """Leading dims flattened: [..., C] contiguous -> (rows, C, ld)."""
    assert x.is_contiguous() and x.dtype == torch.float32
    return x.numel() // x.shape[-1], x.shape[-1], x.shape[-1]


def linear(x1, Wt, cout, w_off1=0, x2=None, w_off2=0, bias=None, act=ACT_NONE, clamp1=0.0, add=None, clamp2=0.0):
    """y[..., cout] = clamp2(clamp1(act(x1 @ Wt[w_off1:w_off1+C1] (+ x2 @ Wt[w_off2:...]) + bias)) + add)."""
    rows, c1, ld1 = _rows(x1)
    c2 = ld2 = 0
    if x2 is not None:
        r2, c2, ld2 = _rows(x2)
        assert r2 == rows
    assert Wt.is_contiguous() and Wt.shape[1] == cout
    y = torch.empty(x1.shape[:-1] + (cout,), dtype=torch.float32, device=x1.device)
    if add is not None:
        assert add.is_contiguous() and add.shape[-1] == cout
    nat.check(nat.lib().ssf_linear(nat.ptr(x1), c1, ld1, nat.ptr(x2), c2, ld2, nat.ptr(Wt), cout, w_off1, w_off2,
                                   nat.ptr(bias), rows, cout, act, float(clamp1), nat.ptr(add), cout, float(clamp2),
                                   nat.ptr(y), cout, nat.stream()))
    return y


EPI_STORE, EPI_MAX, EPI_DOT = 0, 1, 2


def set_dense_variant(mode):
    """Kernel-variant policy of ``dense_tc`` for small layers (include/ssf_dense.h ``ssf_dense_set_variant``: 0 one CTA per SM
    everywhere, 1 default policy, 2 light variant wherever eligible); returns the previous mode.  Results are bit-identical."""
    return int(nat.lib().ssf_dense_set_variant(int(mode)))


def set_dense_tma(level):
    """Tensor-map TMA paths of ``dense_tc``: 0 off, 1 (default) plain-row A operand + STORE epilogue, 2 also the gathered rows of
    grouped layers (tile::gather4; measured slower, kept for the record).  Returns the previous level; results are bit-identical."""
    return int(nat.lib().ssf_dense_set_tma(int(level)))


def dense_tc(wimg, N, K, *, x1=None, x2=None, G=None, offG=0, H=None, offH=0, b1=None, Wd1=None, act1=ACT_NONE, idx=None,
             pos_src=None, pos_q=None, bias=None, Hq=None, Wd2=None, act=ACT_NONE, epi=EPI_STORE, wvec=None, b0=0.0, S=0):
    """Tensor-core dense layer (csrc/dense_tc.cu, include/ssf_dense.h).  Rows: plain ``x1 | x2`` ([..., c]) or the grouped
    first layer on the fly (``G`` [B,Nsrc,ldG], ``idx`` [B,Nq,S], optional per-point ``H`` [B,Nq,ldH]).  Returns
    STORE: [rows..., N]; MAX: [B, Nq, N]; DOT: [rows...]."""
    a = nat.DenseArgs()
    p = nat.ptr
    a.K, a.N, a.wimg = K, N, p(wimg)
    a.S = S                      # plain rows pooled in groups of S (MAX epilogue without a gather)
    if idx is not None:
        B, Nq, S = idx.shape
        a.idx, a.S, a.Nq = p(idx), S, Nq
        a.pos_src, a.pos_q = p(pos_src), p(pos_q)
    if G is not None:
        a.a_mode = 1
        a.G, a.ldG, a.offG, a.Nsrc = p(G), G.shape[-1], offG, G.shape[1]
        if H is not None:
            a.H, a.ldH, a.offH = p(H), H.shape[-1], offH
        a.b1, a.Wd1, a.act1 = p(b1), p(Wd1), act1
        rows, lead = B * Nq * S, (B, Nq, S)
    else:
        a.a_mode = 0
        rows, c1, ld1 = _rows(x1)
        a.x1, a.c1, a.ld1 = p(x1), c1, ld1
        if x2 is not None:
            r2, c2, ld2 = _rows(x2)
            assert r2 == rows
            a.x2, a.c2, a.ld2 = p(x2), c2, ld2
        lead = tuple(x1.shape[:-1])
        if pos_src is not None:
            a.Nsrc = pos_src.shape[1]
    a.rows = rows
    a.bias, a.Wd2, a.act, a.epi_mode = p(bias), p(Wd2), act, epi
    if Hq is not None:
        a.Hq, a.ldHq = p(Hq), Hq.shape[-1]
    dev = (x1 if x1 is not None else G).device
    if epi == EPI_STORE:
        y = torch.empty(lead + (N,), dtype=torch.float32, device=dev)
        a.ldy = N
    elif epi == EPI_MAX:
        y = torch.empty((rows // S, N) if idx is None else (idx.shape[0], idx.shape[1], N), dtype=torch.float32, device=dev)
        a.ldy = N
    else:
        y = torch.empty(lead, dtype=torch.float32, device=dev)
        a.wvec, a.b0 = p(wvec), float(b0)
    a.y = p(y)
    nat.check(nat.lib().ssf_dense_tc(ctypes.byref(a), nat.stream()))
    return y


def attention_mix(A, Aw):
    """A, Aw [..., 16, m] -> (A + Q.Aw, Aw + Q^T.A)  (ASF/utils/soflow.py:420-422,453-458)."""
    m = A.shape[-1]
    n_points = A.numel() // (16 * m)
    Amix, Awmix = torch.empty_like(A), torch.empty_like(Aw)
    nat.check(nat.lib().ssf_attention_mix(nat.ptr(A), nat.ptr(Aw), n_points, m, nat.ptr(Amix), nat.ptr(Awmix), nat.stream()))
    return Amix, Awmix


def softmax_pool(g, C):
    """g [B,N1,16], C [B,N1,16,m] -> (cost_fwd [B,N1,m], channel-major copy [B,m,N1])  (ASF/utils/soflow.py:469,486)."""
    B, N1, _, m = C.shape
    out_pm = torch.empty(B, N1, m, dtype=torch.float32, device=C.device)
    out_cm = torch.empty(B, m, N1, dtype=torch.float32, device=C.device)
    nat.check(nat.lib().ssf_softmax_pool(nat.ptr(g), nat.ptr(C), B, N1, m, nat.ptr(out_pm), nat.ptr(out_cm), nat.stream()))
    return out_pm, out_cm


def gather_rows(src, idx):
    """src [B,N,C], idx i32 [B,M] -> [B,M,C]."""
    B, N, C = src.shape
    M = idx.shape[1]
    out = torch.empty(B, M, C, dtype=torch.float32, device=src.device)
    nat.check(nat.lib().ssf_gather_rows(nat.ptr(src), nat.ptr(idx), B, N, M, C, nat.ptr(out), nat.stream()))
    return out


def transpose(x):
    """[B,R,C] -> [B,C,R] (contiguous)."""
    B, R, C = x.shape
    out = torch.empty(B, C, R, dtype=torch.float32, device=x.device)
    nat.check(nat.lib().ssf_transpose(nat.ptr(x), B, R, C, nat.ptr(out), nat.stream()))
    return out


def fps(xyz, npoint):
    B, N, _ = xyz.shape
    out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    nat.check(nat.lib().ssf_furthest_point_sample(nat.ptr(xyz), B, N, npoint, nat.ptr(out), nat.stream()))
    return out


# reference clouds at least this large go through the Hilbert-ordered block search (csrc/knn_blocks.cu); smaller ones through the
# brute-force scan.  Both return identical indices.
KNN_BLOCKS_MIN_REF = 256   # (measured at 128 clouds: 512 queries in 256 points 0.098 ms scanned, 0.064 ms with index build + search)
KNN_BLOCKS_MAX_REF = 131072   # one-level index up to 16384 points, two-level (super-blocks) above
_knn_tls = threading.local()   # per-thread: forwards running in different Python threads (or DataParallel replicas) never share it


def _knn_cache():
    c = getattr(_knn_tls, "cache", None)
    if c is None:
        c = _knn_tls.cache = {}
    return c


def knn_cache_clear():
    """Drops the calling thread's per-cloud search structures (call at the start of every forward: clouds are new tensors each
    time)."""
    _knn_cache().clear()


def _knn_blocks(ref):
    key = (ref.device.index, ref.data_ptr(), tuple(ref.shape), tuple(ref.stride()))
    _knn_cache_ = _knn_cache()
    hit = _knn_cache_.get(key)
    if hit is None:
        B, Nr, _ = ref.shape
        ws = torch.empty(int(nat.lib().ssf_knn_blocks_workspace_floats(B, Nr)), dtype=torch.float32, device=ref.device)
        nat.check(nat.lib().ssf_knn_blocks_build(nat.ptr(ref), B, Nr, nat.ptr(ws), nat.stream()))
        hit = (ref, ws)  # keeping `ref` alive pins its storage, so the pointer in the key cannot be recycled
        _knn_cache_[key] = hit
    return hit[1]


def knn_idx(k, query, ref, offset=None):
    """Indices only (every hot call site discards the distances): i32 [B,Nq,k]."""
    B, Nq, _ = query.shape
    Nr = ref.shape[1]
    idx = torch.empty(B, Nq, k, dtype=torch.int32, device=query.device)
    if KNN_BLOCKS_MIN_REF <= Nr <= KNN_BLOCKS_MAX_REF:
        ws = _knn_blocks(ref)
        nat.check(nat.lib().ssf_knn_blocks_search(k, nat.ptr(query), nat.ptr(offset), nat.ptr(ws), B, Nq, Nr, None, nat.ptr(idx),
                                                  nat.stream()))
        return idx
    nat.check(nat.lib().ssf_knn_offset(k, nat.ptr(query), nat.ptr(offset), nat.ptr(ref), B, Nq, Nr, None, nat.ptr(idx),
                                       nat.stream()))
    return idx


def interpolate(query, src_pos, src_val, idx, mode=0, clampv=100.0, k=None):
    """idx i32 [B,N,kk] with kk >= k: the first k entries of every row are used (default k = kk)."""
    B, N, _ = query.shape
    M, C = src_val.shape[1], src_val.shape[2]
    ld = idx.shape[2]
    k = ld if k is None else k
    out = torch.empty(B, N, C, dtype=torch.float32, device=query.device)
    nat.check(nat.lib().ssf_interpolate(nat.ptr(query), nat.ptr(src_pos), nat.ptr(src_val), nat.ptr(idx), ld, B, N, M, C, k,
                                        mode, float(clampv), nat.ptr(out), nat.stream()))
    return out


def group_mlp_max(G, idx, pos_src, pos_q, Wd, bias1, W2t, b2, C2, W3t=None, b3=None, C3=0, H=None, act=ACT_RELU):
    """rows (n,s): x = act(G[idx] + H[n] + Wd.(pos_src[idx]-pos_q[n]) + bias1) -> dense layers -> max over s."""
    B, Nsrc, C1 = G.shape
    Nq, S = idx.shape[1], idx.shape[2]
    clast = C3 if C3 > 0 else C2
    out = torch.empty(B, Nq, clast, dtype=torch.float32, device=G.device)
    nat.check(nat.lib().ssf_group_mlp_max(nat.ptr(G), nat.ptr(H), nat.ptr(bias1), nat.ptr(Wd), nat.ptr(pos_src),
                                          nat.ptr(pos_q), nat.ptr(idx), nat.ptr(W2t), nat.ptr(b2), C2, nat.ptr(W3t),
                                          nat.ptr(b3), C3, B, Nsrc, Nq, S, C1, act, nat.ptr(out), nat.stream()))
    return out


def cost_volume(Gab, Hab, w, H3, xyz1, xyz2, idx, idxw, m):
    """Fused cost-volume core; ``w`` is the per-level dict from prepare_weights.  Returns
    (cost_fwd [B,N1,m], cost_fwd_cm [B,m,N1], gw [B,N1*16], Cw [B,N1*16,m])."""
    B, N1, _ = xyz1.shape
    N2 = xyz2.shape[1]
    dev = xyz1.device
    cost_fwd = torch.empty(B, N1, m, dtype=torch.float32, device=dev)
    cost_fwd_cm = torch.empty(B, m, N1, dtype=torch.float32, device=dev)
    gw = torch.empty(B, N1 * 16, dtype=torch.float32, device=dev)
    Cw = torch.empty(B, N1 * 16, m, dtype=torch.float32, device=dev)
    p = nat.ptr
    if USE_TC and "tc_blob" in w:
        nat.check(nat.lib().ssf_cost_volume_tc(p(Gab), p(Hab), p(H3), p(w["tc_blob"]), p(w["tc_par"]), p(xyz1), p(xyz2), p(idx),
                                               p(idxw), B, N1, N2, m, p(cost_fwd), p(cost_fwd_cm), p(gw), p(Cw), 0, nat.stream()))
        return cost_fwd, cost_fwd_cm, gw, Cw
    nat.check(nat.lib().ssf_cost_volume(p(Gab), p(Hab), p(w["W2a"]), p(w["b2a"]), p(w["W2w"]), p(w["b2w"]), p(w["W3a"]),
                                        p(H3), p(w["W3d"]), p(w["W3b"]), p(w["b3b"]), p(w["Wn1"]), p(w["bn1"]), p(w["Wn2"]),
                                        p(w["bn2"]), p(w["wn3"]), float(w["bn3"]), p(xyz1), p(xyz2), p(idx), p(idxw), B, N1,
                                        N2, m, p(cost_fwd), p(cost_fwd_cm), p(gw), p(Cw), nat.stream()))
    return cost_fwd, cost_fwd_cm, gw, Cw


def build_csr(key, n_seg):
    """key i32/i64 [B,L] -> workspace tensor describing, per segment, its rows in ascending order."""
    B, L = key.shape
    ws = torch.empty(int(nat.lib().ssf_csr_workspace_ints(B, L, n_seg)), dtype=torch.int32, device=key.device)
    fn = nat.lib().ssf_build_csr_i64 if key.dtype == torch.int64 else nat.lib().ssf_build_csr_i32
    nat.check(fn(nat.ptr(key), B, L, n_seg, nat.ptr(ws), nat.stream()))
    return ws


def segment_softmax_sum(logit, val, csr, n_seg):
    """logit [B,L], val [B,L,C] -> [B,n_seg,C] (zeros for empty segments)."""
    B, L, C = val.shape
    out = torch.empty(B, n_seg, C, dtype=torch.float32, device=val.device)
    nat.check(nat.lib().ssf_segment_softmax_sum(nat.ptr(logit), nat.ptr(val), nat.ptr(csr), B, L, C, n_seg, nat.ptr(out),
                                                nat.stream()))
    return out


def frontend(points, flow, mode=1, in_mask=None, sem=None, movable=(), inst=None, n_inst=0, tau=0.10, want_pose=False):
    """points, flow f32 [B,N,3] -> (mask u8 [B,N], odom f64 [B,7] (, pose f64 [B,12]))."""
    B, N, _ = points.shape
    dev = points.device
    mask = torch.empty(B, N, dtype=torch.uint8, device=dev)
    odom = torch.empty(B, 7, dtype=torch.float64, device=dev)
    pose = torch.empty(B, 12, dtype=torch.float64, device=dev) if want_pose else None
    bits = 0
    for c in movable:
        if not 0 <= int(c) < 64:
            raise nat.SsfError("movable class ids must lie in [0, 64) (they are passed to the kernel as a 64-bit set); got %d" % int(c))
        bits |= 1 << int(c)
    nat.check(nat.lib().ssf_frontend(nat.ptr(points), nat.ptr(flow), B, N, mode, nat.ptr(in_mask), nat.ptr(sem), bits,
                                     nat.ptr(inst), int(n_inst), float(tau), nat.ptr(mask), nat.ptr(odom), nat.ptr(pose),
                                     nat.stream()))
    return (mask, odom, pose) if want_pose else (mask, odom)


def solve_rt_f64(src, dst):
    """src, dst f64 [B,M,3] -> (odom f64 [B,7], pose f64 [B,12]): dst ~= R src + t, reduced in float64."""
    B, M, _ = src.shape
    odom = torch.empty(B, 7, dtype=torch.float64, device=src.device)
    pose = torch.empty(B, 12, dtype=torch.float64, device=src.device)
    nat.check(nat.lib().ssf_solve_rt_f64(nat.ptr(src), nat.ptr(dst), B, M, nat.ptr(odom), nat.ptr(pose), nat.stream()))
    return odom, pose


def gmm_mask(points, flow, max_iter=100, tol=1e-3, want_info=False):
    """The reference's noSeg masker (2-component GMM on [flow | xyz], majority = background; csrc/gmm.cu).
    points, flow f32 [B,N,3] -> mask u8 [B,N] (0 = background) (, info f64 [B,4] = n_iter, lower bound, converged, bg count)."""
    B, N, _ = points.shape
    mask = torch.empty(B, N, dtype=torch.uint8, device=points.device)
    info = torch.empty(B, 4, dtype=torch.float64, device=points.device) if want_info else None
    nat.check(nat.lib().ssf_gmm_mask(nat.ptr(points), nat.ptr(flow), B, N, int(max_iter), float(tol), nat.ptr(mask), nat.ptr(info),
                                     nat.stream()))
    return (mask, info) if want_info else mask
