"""Out-of-bounds WRITE check without compute-sanitizer (the tool is closed on this GPU pool, profiles/r02_sanitizer_refused.log):
every output buffer handed to the C ABI sits between two canary regions filled with a bit pattern; after the launch the canaries
must be untouched and the payload must equal the result of the ordinary (torch-allocated) call.  Ragged sizes on purpose: last
tiles, partial warps, rows past the end of a TMA box."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CANARY = 2048  # bytes on each side


class Guarded:
    """A CUDA buffer of `shape` / `dtype` with canaries before and after it."""

    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), dtype
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        pad = (-n) % 256
        self.raw = torch.full((CANARY + n + pad + CANARY,), 0xA5, dtype=torch.uint8, device="cuda")
        self.view = self.raw[CANARY:CANARY + n].view(dtype).view(self.shape)
        self.n, self.pad = n, pad

    def check(self):
        a = self.raw[:CANARY]
        b = self.raw[CANARY + self.n:]
        assert bool((a == 0xA5).all()), "write BEFORE the buffer"
        assert bool((b == 0xA5).all()), "write PAST the buffer"


def _check(rc):
    from ssf_slam_b200 import _native as nat
    nat.check(rc)
    torch.cuda.synchronize()


def test_point_operators_write_inside_their_outputs():
    from ssf_slam_b200 import _native as nat, pointnet2_utils as pu
    L, p, st = nat.lib(), nat.ptr, nat.stream
    g = torch.Generator().manual_seed(1)
    for B, N, M, k in ((3, 5003, 777, 16), (2, 20001, 1031, 7), (1, 70001, 513, 16)):
        xyz = (torch.randn(B, N, 3, generator=g) * torch.tensor([30.0, 20.0, 2.0])).cuda()
        q = xyz[:, :M].contiguous()
        # FPS (block / cluster kernels)
        out = Guarded((B, M), torch.int32)
        _check(L.ssf_furthest_point_sample(p(xyz), B, N, M, p(out.view), st()))
        out.check()
        assert torch.equal(out.view, pu.furthest_point_sample(xyz, M))
        # block index build + search (one level / two levels) and the scan
        ws = Guarded((int(L.ssf_knn_blocks_workspace_floats(B, N)),), torch.float32)
        _check(L.ssf_knn_blocks_build(p(xyz), B, N, p(ws.view), st()))
        ws.check()
        d, i = Guarded((B, M, k), torch.float32), Guarded((B, M, k), torch.int32)
        _check(L.ssf_knn_blocks_search(k, p(q), None, p(ws.view), B, M, N, p(d.view), p(i.view), st()))
        d.check(), i.check()
        d2, i2 = Guarded((B, M, k), torch.float32), Guarded((B, M, k), torch.int32)
        _check(L.ssf_knn_offset(k, p(q), None, p(xyz), B, M, N, p(d2.view), p(i2.view), st()))
        d2.check(), i2.check()
        assert torch.equal(i.view, i2.view) and torch.equal(d.view, d2.view)
        # ball query: scan and index
        for fn_idx in (0, 1):
            bi, bc = Guarded((B, M, 16), torch.int32), Guarded((B, M), torch.int32)
            if fn_idx == 0:
                _check(L.ssf_ball_query(1.5, 16, p(xyz), p(q), B, N, M, p(bi.view), p(bc.view), st()))
            else:
                _check(L.ssf_ball_query_blocks(1.5, 16, p(q), p(ws.view), B, N, M, p(bi.view), p(bc.view), st()))
            bi.check(), bc.check()
            want_i, want_c = pu.ball_query(1.5, 16, xyz, q, return_count=True, use_index=False)
            assert torch.equal(bi.view, want_i) and torch.equal(bc.view, want_c)
        # gathers (staged and plain paths) on ragged channel counts
        C = 5
        feat = torch.randn(B, C, N, generator=g).cuda()
        grp = Guarded((B, C, M, k), torch.float32)
        _check(L.ssf_grouping_operation(p(feat), p(i.view), B, C, N, M, k, p(grp.view), st()))
        grp.check()
        assert torch.equal(grp.view, pu.grouping_operation(feat, i.view))
        w3 = torch.rand(B, M, 3, generator=g).cuda()
        i3 = i.view[:, :, :3].contiguous()
        ti = Guarded((B, C, M), torch.float32)
        _check(L.ssf_three_interpolate(p(feat), p(i3), p(w3), B, C, N, M, p(ti.view), st()))
        ti.check()
        assert torch.equal(ti.view, pu.three_interpolate(feat, i3, w3))


def test_segmented_ops_and_frontend_write_inside_their_outputs():
    from ssf_slam_b200 import _native as nat, functional as F_
    L, p, st = nat.lib(), nat.ptr, nat.stream
    g = torch.Generator().manual_seed(2)
    B, Lr, C, n_seg = 2, 5003, 64, 301
    key = torch.randint(0, n_seg, (B, Lr), generator=g, dtype=torch.int32).cuda()
    key[:, :2000] = 7                                   # one long segment (warp-cooperative sort path)
    ws = Guarded((int(L.ssf_csr_workspace_ints(B, Lr, n_seg)),), torch.int32)
    _check(L.ssf_build_csr_i32(p(key), B, Lr, n_seg, p(ws.view), st()))
    ws.check()
    logit, val = torch.randn(B, Lr, generator=g).cuda(), torch.randn(B, Lr, C, generator=g).cuda()
    out = Guarded((B, n_seg, C), torch.float32)
    _check(L.ssf_segment_softmax_sum(p(logit), p(val), p(ws.view), B, Lr, C, n_seg, p(out.view), st()))
    out.check()
    assert torch.equal(out.view, F_.segment_softmax_sum(logit, val, F_.build_csr(key, n_seg), n_seg))
    # mask + pose, GMM masker, fp64 pose
    N = 4099
    pts = (torch.randn(B, N, 3, generator=g) * 20).cuda()
    flow = (0.3 + 0.05 * torch.randn(B, N, 3, generator=g)).cuda()
    mask, odom, pose = Guarded((B, N), torch.uint8), Guarded((B, 7), torch.float64), Guarded((B, 12), torch.float64)
    _check(L.ssf_frontend(p(pts), p(flow), B, N, 1, None, None, 0, None, 0, 0.1, p(mask.view), p(odom.view), p(pose.view), st()))
    mask.check(), odom.check(), pose.check()
    m2, o2 = F_.frontend(pts, flow, mode=1, tau=0.1)
    assert torch.equal(mask.view, m2) and torch.equal(odom.view, o2)
    gm, info = Guarded((B, N), torch.uint8), Guarded((B, 4), torch.float64)
    _check(L.ssf_gmm_mask(p(pts), p(flow), B, N, 100, 1e-3, p(gm.view), p(info.view), st()))
    gm.check(), info.check()
    assert torch.equal(gm.view, F_.gmm_mask(pts, flow))


@pytest.mark.parametrize("tma", [1, 0])
def test_tensor_core_layers_write_inside_their_outputs(tma):
    """dense_tc (plain rows with and without the tensor-map TMA paths, grouped + max) and the fused cost volume with ragged row /
    point counts: the TMA store must clip at the last row, the epilogues must not write past it."""
    from ssf_slam_b200 import _native as nat, functional as F_, tc
    from ssf_slam_b200.model import prepare_weights
    from ssf_slam_b200.weights import random_init_state_dict
    L, p, st = nat.lib(), nat.ptr, nat.stream
    g = torch.Generator().manual_seed(3)
    prev = F_.set_dense_tma(tma)
    try:
        for rows, K, N in ((1000 + 37, 64, 64), (148 * 128 + 77, 96, 128), (300, 256, 256)):
            X = torch.randn(rows, K, generator=g).cuda()
            img = tc.dense_image(torch.randn(N, K, generator=g) / K ** 0.5).cuda()
            bias = torch.randn(N, generator=g).cuda()
            want = F_.dense_tc(img, N, K, x1=X, bias=bias, act=2)
            y = Guarded((rows, N), torch.float32)
            a = nat.DenseArgs()
            a.K, a.N, a.wimg, a.a_mode, a.rows = K, N, p(img), 0, rows
            a.x1, a.c1, a.ld1 = p(X), K, K
            a.bias, a.act, a.epi_mode, a.y, a.ldy = p(bias), 2, 0, p(y.view), N
            _check(L.ssf_dense_tc(ctypes.byref(a), st()))
            y.check()
            assert torch.equal(y.view, want)
    finally:
        F_.set_dense_tma(prev)
    # fused cost volume, N1 not a multiple of the 8-point tile
    W = prepare_weights(random_init_state_dict(0), torch.device("cuda:0"))["flow0_r"]
    B, N1, N2, m = 2, 1003, 900, 64
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    Gab, Hab, H3 = r(B, N2, 128) * 0.7, r(B, N1, 128) * 0.7, r(B, N1, 64) * 0.5
    x1, x2 = r(B, N1, 3) * 10, r(B, N2, 3) * 10
    idx = torch.randint(0, N2, (B, N1, 16), generator=g, dtype=torch.int32).cuda()
    idxw = torch.randint(0, N2, (B, N1, 16), generator=g, dtype=torch.int32).cuda()
    want = F_.cost_volume(Gab, Hab, W, H3, x1, x2, idx, idxw, m)
    outs = [Guarded((B, N1, m), torch.float32), Guarded((B, m, N1), torch.float32), Guarded((B, N1 * 16), torch.float32),
            Guarded((B, N1 * 16, m), torch.float32)]
    _check(L.ssf_cost_volume_tc(p(Gab), p(Hab), p(H3), p(W["tc_blob"]), p(W["tc_par"]), p(x1), p(x2), p(idx), p(idxw), B, N1, N2, m,
                                p(outs[0].view), p(outs[1].view), p(outs[2].view), p(outs[3].view), 0, st()))
    for o, w_ in zip(outs, want):
        o.check()
        assert torch.equal(o.view, w_)
