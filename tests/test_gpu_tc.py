"""GPU: tcgen05 kind::tf32 bring-up -- 3xTF32 GEMM against fp64, A operand from TMEM and from shared memory."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(K, N, mode, passes, seed=0):
    from ssf_slam_b200 import _native as nat
    from ssf_slam_b200 import tc
    nat.require_device()
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(128, K, generator=g)
    W = torch.randn(N, K, generator=g)
    hi, lo = tc.weight_image(W)
    Y = torch.full((128, N), float("nan"), device="cuda")
    Xd, hid, lod = X.cuda(), hi.cuda(), lo.cuda()  # keep the device tensors alive across the launch
    nat.check(nat.dev_lib().ssf_tc_gemm_test(nat.ptr(Xd), nat.ptr(hid), nat.ptr(lod), K, N, mode, passes, nat.ptr(Y), nat.stream()), nat.dev_lib())
    torch.cuda.synchronize()
    ref = (X.double() @ W.double().t())
    err = float((Y.cpu().double() - ref).abs().max())
    scale = float(ref.abs().max())
    return err / scale


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("K,N", [(64, 64), (64, 32), (64, 128), (32, 64), (128, 64), (64, 256), (8, 16)])
def test_tc_gemm_3xtf32(mode, K, N):
    rel = _run(K, N, mode, 3)
    print("mode %d K %d N %d: 3xTF32 max rel err %.3g" % (mode, K, N, rel))
    assert rel < 2e-6


@pytest.mark.parametrize("mode", [0, 1])
def test_tc_gemm_plain_tf32_is_coarser(mode):
    rel1 = _run(64, 64, mode, 1)
    rel3 = _run(64, 64, mode, 3)
    print("mode %d: 1xTF32 %.3g, 3xTF32 %.3g" % (mode, rel1, rel3))
    assert 1e-5 < rel1 < 5e-3 and rel3 < rel1 / 50


# ------------------------------------------------------------------ tensor-core cost volume vs the SIMT fp32 kernel

def _cv_inputs(B, N1, N2, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    Gab, Hab, H3 = r(B, N2, 128) * 0.7, r(B, N1, 128) * 0.7, r(B, N1, 64) * 0.5
    xyz1, xyz2 = r(B, N1, 3) * 3, r(B, N2, 3) * 3
    idx = torch.randint(0, N2, (B, N1, 16), generator=g, dtype=torch.int32)
    idxw = torch.randint(0, N2, (B, N1, 16), generator=g, dtype=torch.int32)
    return [t.cuda().contiguous() for t in (Gab, Hab, H3, xyz1, xyz2, idx, idxw)]


@pytest.fixture(scope="module")
def cv_weights():
    from oracle import tflow_port as tp  # checker-side weight generator only
    from ssf_slam_b200.model import prepare_weights
    return prepare_weights(tp.random_init_state_dict(0), torch.device("cuda:0"))


@pytest.mark.parametrize("level,B,N1,N2,n_sm", [("flow1_r", 1, 8, 40, 0), ("flow1_r", 2, 203, 150, 0), ("flow0_r", 3, 2048, 1024, 0),
                                                ("flow0_r", 2, 1024, 512, 3)])
def test_cost_volume_tc_matches_simt(cv_weights, level, B, N1, N2, n_sm):
    from ssf_slam_b200 import functional as F_
    from ssf_slam_b200 import _native as nat
    w = cv_weights[level]
    Gab, Hab, H3, xyz1, xyz2, idx, idxw = _cv_inputs(B, N1, N2, 7)
    F_.USE_TC = False
    try:
        ref = F_.cost_volume(Gab, Hab, w, H3, xyz1, xyz2, idx, idxw, 64)
    finally:
        F_.USE_TC = True
    if n_sm == 0:
        got = F_.cost_volume(Gab, Hab, w, H3, xyz1, xyz2, idx, idxw, 64)
    else:  # few CTAs -> many tiles per persistent CTA (exercises the barrier phase bookkeeping)
        got = [torch.full_like(t, float("nan")) for t in ref]
        p = nat.ptr
        nat.check(nat.lib().ssf_cost_volume_tc(p(Gab), p(Hab), p(H3), p(w["tc_blob"]), p(w["tc_par"]), p(xyz1), p(xyz2), p(idx),
                                               p(idxw), B, N1, N2, 64, p(got[0]), p(got[1]), p(got[2]), p(got[3]), n_sm,
                                               nat.stream()))
    torch.cuda.synchronize()
    for name, r_, g_ in zip(("cost_fwd", "cost_fwd_cm", "gw", "Cw"), ref, got):
        scale = max(1.0, float(r_.abs().max()))
        err = float((r_ - g_).abs().max())
        print(level, name, "max-abs diff %.3g (scale %.3g)" % (err, scale))
        assert err < 2e-5 * scale, (name, err, scale)


# ------------------------------------------------------------------ generic tensor-core dense layer

def _act(x, act):
    if act == 1:
        return x.clamp(min=0)
    if act == 2:
        return torch.where(x > 0, x, 0.1 * x)
    return x


@pytest.mark.parametrize("rows,K,N,two_seg,act", [(128, 32, 32, False, 0), (1000, 96, 128, True, 2), (4096, 256, 256, False, 1),
                                                  (300, 512, 256, True, 1), (2048, 256, 512, False, 1), (257, 64, 64, False, 2),
                                                  # many tiles per persistent CTA: barrier phase bookkeeping, both producer warpgroups
                                                  (148 * 128 * 5 + 77, 64, 64, False, 2), (148 * 128 * 3, 32, 32, False, 1),
                                                  (148 * 128 * 4 + 1, 256, 256, False, 1), (148 * 128 * 3 + 130, 96, 128, True, 0),
                                                  (148 * 128 * 2 + 5, 256, 128, False, 2)])
def test_dense_tc_rows(rows, K, N, two_seg, act):
    from ssf_slam_b200 import functional as F_, tc
    g = torch.Generator().manual_seed(rows + K + N)
    X = torch.randn(rows, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    img = tc.dense_image(W).cuda()
    if two_seg:
        c1 = 64 if K > 64 else 32
        y = F_.dense_tc(img, N, K, x1=X[:, :c1].contiguous().cuda(), x2=X[:, c1:].contiguous().cuda(), bias=b.cuda(), act=act)
    else:
        y = F_.dense_tc(img, N, K, x1=X.cuda(), bias=b.cuda(), act=act)
    ref = _act(X.double() @ W.double().t() + b.double(), act)
    err = float((y.cpu().double() - ref).abs().max())
    print("dense_tc rows", rows, K, N, "max-abs err %.3g" % err)
    assert err < 3e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,Nsrc,Nq,S,K,N,epi", [(2, 300, 100, 16, 64, 64, 0), (3, 200, 72, 8, 256, 512, 1), (2, 500, 333, 16, 128, 128, 1),
                                                  (2, 128, 64, 16, 128, 64, 2), (3, 2048, 4000, 16, 64, 64, 1), (2, 512, 5000, 8, 32, 64, 0)])
def test_dense_tc_grouped(B, Nsrc, Nq, S, K, N, epi):
    """Grouped first layer on the fly (gather + per-point block + direction term), per-point epilogue add, max / dot."""
    from ssf_slam_b200 import functional as F_, tc
    g = torch.Generator().manual_seed(B * 1000 + Nq)
    r = lambda *s: torch.randn(*s, generator=g)
    G, H, b1, Wd1 = r(B, Nsrc, K + 32), r(B, Nq, K), r(K), r(3, K) * 0.3
    ps, pq = r(B, Nsrc, 3), r(B, Nq, 3)
    idx = torch.randint(0, Nsrc, (B, Nq, S), generator=g, dtype=torch.int32)
    W, bias, Hq, Wd2, wvec = r(N, K) / K ** 0.5, r(N), r(B, Nq, N), r(3, N) * 0.3, r(N)
    li = idx.long()
    bi = torch.arange(B)[:, None, None]
    dirs = (ps[bi, li] - pq[:, :, None, :]).double()                              # [B,Nq,S,3]
    A = _act(G[bi, li][..., 32:].double() + H[:, :, None, :].double() + b1.double() + dirs @ Wd1.double(), 2)
    D = A @ W.double().t() + bias.double() + Hq[:, :, None, :].double() + dirs @ Wd2.double()
    D = _act(D, 1)
    cu = lambda t: t.cuda().contiguous()
    y = F_.dense_tc(cu(tc.dense_image(W)), N, K, G=cu(G), offG=32, H=cu(H), b1=cu(b1), Wd1=cu(Wd1), act1=2, idx=cu(idx),
                    pos_src=cu(ps), pos_q=cu(pq), bias=cu(bias), Hq=cu(Hq), Wd2=cu(Wd2), act=1, epi=epi, wvec=cu(wvec), b0=0.25)
    ref = D if epi == 0 else (D.max(dim=2).values if epi == 1 else D @ wvec.double() + 0.25)
    err = float((y.cpu().double() - ref).abs().max())
    print("dense_tc grouped epi", epi, "max-abs err %.3g" % err, "scale %.3g" % float(ref.abs().max()))
    assert y.shape == ref.shape and err < 3e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("rows,K,N,grouped,epi", [(148 * 128 * 7 + 33, 64, 64, False, 0), (148 * 128 * 5, 96, 64, False, 0),
                                                  (148 * 128 * 3 + 1, 128, 32, False, 0), (148 * 128 * 9 + 16, 64, 64, True, 1),
                                                  (148 * 128 * 4 + 64, 32, 64, True, 1), (128 * 40, 192, 64, False, 0),
                                                  (148 * 128 * 6 + 48, 128, 64, True, 2), (100, 32, 32, False, 0),
                                                  # 256-column tiles: the two-epilogue-warpgroup variant
                                                  (148 * 128 * 2 + 77, 256, 256, False, 0), (148 * 128 + 16, 128, 256, True, 1),
                                                  (128 * 9, 256, 512, False, 0)])
def test_dense_tc_light_variant_is_bit_identical(rows, K, N, grouped, epi):
    """The kernel variants (two CTAs per SM for N <= 64, two epilogue warpgroups for 256-column tiles) must reproduce the base
    variant bit for bit."""
    from ssf_slam_b200 import functional as F_, tc
    g = torch.Generator().manual_seed(rows + K + N)
    r = lambda *s: torch.randn(*s, generator=g)
    W, bias = r(N, K) / K ** 0.5, r(N)
    img = tc.dense_image(W).cuda()
    if grouped:
        S, Nsrc = 16, 777
        Nq = rows // S
        G, H, b1, Wd1 = r(1, Nsrc, K).cuda(), r(1, Nq, K).cuda(), r(K).cuda(), (r(3, K) * 0.3).cuda()
        ps, pq = r(1, Nsrc, 3).cuda(), r(1, Nq, 3).cuda()
        idx = torch.randint(0, Nsrc, (1, Nq, S), generator=g, dtype=torch.int32).cuda()
        wv = r(N).cuda()
        run = lambda: F_.dense_tc(img, N, K, G=G, H=H, b1=b1, Wd1=Wd1, act1=2, idx=idx, pos_src=ps, pos_q=pq, bias=bias.cuda(),
                                  act=1, epi=epi, wvec=wv, b0=0.5)
    else:
        X = r(rows, K).cuda()
        run = lambda: F_.dense_tc(img, N, K, x1=X, bias=bias.cuda(), act=2)
    prev = F_.set_dense_variant(2)
    try:
        y_light = run()
        F_.set_dense_variant(1)      # the default policy (wide / full / light chosen per layer)
        y_default = run()
        F_.set_dense_variant(0)
        y_heavy = run()
    finally:
        F_.set_dense_variant(prev)
    assert torch.equal(y_light, y_heavy) and torch.equal(y_default, y_heavy)
    if not grouped:
        ref = _act(X.cpu().double() @ W.double().t() + bias.double(), 2)
    else:   # MAX / DOT without a per-row epilogue term: bias and activation are applied after the pooling
        li = idx.cpu().long()[0]
        dirs = (ps.cpu()[0][li] - pq.cpu()[0][:, None, :]).double()
        A = _act(G.cpu()[0][li].double() + H.cpu()[0][:, None, :].double() + b1.cpu().double() + dirs @ Wd1.cpu().double(), 2)
        D = _act(A @ W.double().t() + bias.double(), 1)
        ref = (D.max(dim=1).values if epi == 1 else D @ wv.cpu().double() + 0.5)[None]
    assert y_light.shape == ref.shape
    assert float((y_light.cpu().double() - ref).abs().max()) < 3e-6 * max(1.0, float(ref.abs().max()))


def test_dense_tc_full_variant_grouped_without_point_block():
    """Grouped pooling layer without a per-point block H (the SA / SU layers): runs the 2 + 2 warpgroup variant by default; must
    equal the base variant bit for bit and the fp64 definition."""
    from ssf_slam_b200 import functional as F_, tc
    g = torch.Generator().manual_seed(5)
    r = lambda *s: torch.randn(*s, generator=g)
    B, Nsrc, Nq, S, K, N = 2, 900, 148 * 8 * 3 + 5, 16, 64, 64
    G, b1, Wd1, ps, pq = r(B, Nsrc, K), r(K), r(3, K) * 0.3, r(B, Nsrc, 3), r(B, Nq, 3)
    idx = torch.randint(0, Nsrc, (B, Nq, S), generator=g, dtype=torch.int32)
    W, bias = r(N, K) / K ** 0.5, r(N)
    cu = lambda t: t.cuda().contiguous()
    img = cu(tc.dense_image(W))
    args = dict(G=cu(G), b1=cu(b1), Wd1=cu(Wd1), act1=1, idx=cu(idx), pos_src=cu(ps), pos_q=cu(pq), bias=cu(bias), act=1, epi=1)
    prev = F_.set_dense_variant(1)
    try:
        y1 = F_.dense_tc(img, N, K, **args)
        F_.set_dense_variant(0)
        y0 = F_.dense_tc(img, N, K, **args)
    finally:
        F_.set_dense_variant(prev)
    assert torch.equal(y1, y0)
    li, bi = idx.long(), torch.arange(B)[:, None, None]
    dirs = (ps[bi, li] - pq[:, :, None, :]).double()
    A = _act(G[bi, li].double() + b1.double() + dirs @ Wd1.double(), 1)
    ref = _act(A @ W.double().t() + bias.double(), 1).max(dim=2).values
    assert float((y1.cpu().double() - ref).abs().max()) < 3e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("rows,K,N,two_seg,epi", [(128, 32, 32, False, 0), (1000, 96, 128, True, 0), (300, 512, 256, True, 0), (257, 64, 64, False, 0),
                                                  (148 * 128 * 3 + 130, 96, 128, True, 0), (148 * 128 * 2 + 5, 256, 256, False, 0),
                                                  (148 * 128 * 2 + 37, 64, 64, False, 0), (4096 * 16, 64, 128, False, 1), (33, 64, 64, False, 0)])
def test_dense_tc_tensor_map_tma_is_bit_identical(rows, K, N, two_seg, epi):
    """The tensor-map TMA paths (cp.async.bulk.tensor loads of the plain-row A operand with zero fill past the last row, bulk
    tensor stores of the STORE epilogue with clipping) against the cp.async / st.global paths: bit-identical, ragged row counts and
    two-segment inputs included; and against the fp64 product."""
    from ssf_slam_b200 import functional as F_, tc
    g = torch.Generator().manual_seed(7 * rows + K + N)
    X = torch.randn(rows, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    img = tc.dense_image(W).cuda()
    kw = dict(bias=b.cuda(), act=2, epi=epi, S=16 if epi == 1 else 0)
    if two_seg:
        c1 = 64 if K > 64 else 32
        xs = dict(x1=X[:, :c1].contiguous().cuda(), x2=X[:, c1:].contiguous().cuda())
    else:
        xs = dict(x1=X.cuda())
    outs = {}
    prev = F_.set_dense_tma(1)
    try:
        for on in (True, False):
            F_.set_dense_tma(1 if on else 0)
            outs[on] = F_.dense_tc(img, N, K, **xs, **kw).cpu()
    finally:
        F_.set_dense_tma(prev)
    assert torch.equal(outs[True], outs[False])
    ref = _act(X.double() @ W.double().t() + b.double(), 2)
    if epi == 1:
        ref = ref.view(rows // 16, 16, N).max(dim=1)[0]
    assert float((outs[True].double() - ref).abs().max()) < (3e-6 if K <= 256 else 6e-6) * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,Nsrc,Nq,S,K,N,epi", [(2, 300, 100, 16, 64, 64, 0), (3, 2048, 4000, 16, 64, 64, 1), (2, 512, 5000, 8, 32, 64, 0),
                                                  (2, 500, 333, 16, 128, 128, 1), (4, 2048, 8192, 16, 64, 64, 1)])
def test_dense_tc_gather4_tma_is_bit_identical(B, Nsrc, Nq, S, K, N, epi):
    """Grouped first layer: the neighbour rows fetched by tensor-map TMA (cp.async.bulk.tensor tile::gather4, four rows per
    instruction, swizzled tiles) against the per-lane cp.async gather: bit-identical outputs (STORE and MAX epilogues, a source
    array wider than K with a column offset, ragged last tile)."""
    from ssf_slam_b200 import functional as F_, tc
    g = torch.Generator().manual_seed(B * 977 + Nq)
    r = lambda *s: torch.randn(*s, generator=g)
    G, H, b1, Wd1 = r(B, Nsrc, K + 32).cuda(), r(B, Nq, K).cuda(), r(K).cuda(), (r(3, K) * 0.3).cuda()
    ps, pq = r(B, Nsrc, 3).cuda(), r(B, Nq, 3).cuda()
    idx = torch.randint(0, Nsrc, (B, Nq, S), generator=g, dtype=torch.int32).cuda()
    img = tc.dense_image(r(N, K) / K ** 0.5).cuda()
    bias = r(N).cuda()
    outs = {}
    prev = F_.set_dense_tma(1)
    try:
        for on in (True, False):
            F_.set_dense_tma(2 if on else 0)
            outs[on] = F_.dense_tc(img, N, K, G=G, offG=32, H=H, b1=b1, Wd1=Wd1, act1=2, idx=idx, pos_src=ps, pos_q=pq, bias=bias, act=1,
                                   epi=epi).cpu()
    finally:
        F_.set_dense_tma(prev)
    assert torch.equal(outs[True], outs[False])
