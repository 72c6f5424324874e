"""GPU parity: point operators through the C ABI vs the CPU oracle (bit-exact indices, exact distances)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import point_ops as po  # noqa: E402  (checker only)


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def pu(oracle_c):
    from ssf_slam_b200 import pointnet2_utils
    return pointnet2_utils


def _cloud(seed, B, N, dup=True, scale=30.0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-scale, scale, (B, N, 3)).astype(np.float32)
    if dup and N >= 64:
        x[:, N // 2:N // 2 + N // 8] = x[:, :N // 8]  # exact duplicates -> exact ties
    return x


def test_golden_point_ops(pu, golden_dir):
    g = np.load(os.path.join(golden_dir, "point_ops.npz"))
    xyz, query = _cuda(g["xyz"]), _cuda(g["query"])
    assert np.array_equal(pu.furthest_point_sample(xyz, 256).cpu().numpy(), g["fps256"])
    for name, k in (("knn16", 16), ("knn7", 7)):
        d, i = pu.knn(k, query, xyz)
        assert np.array_equal(i.cpu().numpy(), g[name + "_idx"])
        assert np.array_equal(d.cpu().numpy(), g[name + "_dist"])
    d, i = pu.three_nn(query, xyz)
    assert np.array_equal(i.cpu().numpy(), g["nn3_idx"]) and np.array_equal(d.cpu().numpy(), g["nn3_dist"])
    for r in (0.5, 2.0, 4.0):
        bi, bc = pu.ball_query(r, 16, xyz, query, return_count=True)
        assert np.array_equal(bi.cpu().numpy(), g["ball_r%g_idx" % r])
        assert np.array_equal(bc.cpu().numpy(), g["ball_r%g_cnt" % r])


def test_reference_twins_golden(pu, golden_dir):
    """a10 pin on the GPU: the CUDA operators against outputs of the reference's OWN in-tree pure-torch twins of the
    extension (ASF/utils/utils.py:68-108, ASF/SetCover.py:39-63; written by oracle/gen_golden_point_twins.py from the
    unmodified reference source on tie-free clouds)."""
    g = np.load(os.path.join(golden_dir, "point_twins.npz"))
    assert np.array_equal(pu.furthest_point_sample(_cuda(g["fps_xyz"]), 512).cpu().numpy(), g["fps_idx"])
    q, r = _cuda(g["knn_query"]), _cuda(g["knn_ref"])
    for k in (3, 8, 16):
        d, i = pu.knn(k, q, r)
        assert np.array_equal(i.cpu().numpy(), g["knn%d_idx" % k])
        assert np.allclose(d.cpu().numpy(), g["knn%d_dist" % k], rtol=1e-5, atol=1e-6)
    d, i = pu.three_nn(q, r)
    assert np.array_equal(i.cpu().numpy(), g["knn3_idx"])
    for rad in (0.5, 1.0, 2.0, 4.0):
        bi, bc = pu.ball_query(rad, 16, _cuda(g["ball_xyz"]), _cuda(g["ball_new_xyz"]), return_count=True)
        assert np.array_equal(bi.cpu().numpy(), g["ball_r%g_idx" % rad])
        assert np.array_equal(bc.cpu().numpy(), g["ball_r%g_cnt" % rad])


@pytest.mark.parametrize("N,npoint", [(8192, 2048), (2048, 512), (512, 256), (256, 128), (100, 37), (1, 1), (5000, 300),
                                      (130, 140), (16384, 1024), (20000, 512), (65536, 2048), (4096, 512), (4000, 100), (3500, 64)])
def test_fps(pu, N, npoint):
    B = 3 if N <= 8192 else 2
    x = _cloud(N, B, N)
    got = pu.furthest_point_sample(_cuda(x), npoint).cpu().numpy()
    assert np.array_equal(got, po.c_fps(x, npoint))


@pytest.mark.parametrize("case", ["n2049", "n4097", "all_equal", "planar", "line", "clusters", "select_all", "more_than_all", "lidar", "tiny_far"])
def test_fps_pruned_kernel_edge_cases(pu, case):
    """The spatially pruned sampler (2048 < N <= 8192: Hilbert-ordered rows + box bound, point_ops.cu fps_pruned_kernel) must stay
    bit-identical to the plain scan of the oracle where its machinery is stressed: nearly empty padded rows, zero extents,
    massive ties (lowest original index wins), tight clusters far apart, and sampling every point."""
    rng = np.random.default_rng(abs(hash(case)) % 1000)
    npoint = 256
    if case == "n2049":
        x = _cloud(11, 2, 2049)
    elif case == "n4097":
        x = _cloud(12, 2, 4097)
    elif case == "all_equal":
        x = np.full((2, 3000, 3), 1.25, np.float32)
    elif case == "planar":
        x = _cloud(13, 2, 6000)
        x[..., 2] = 0.5
    elif case == "line":
        x = np.zeros((2, 5000, 3), np.float32)
        x[..., 0] = rng.uniform(-50, 50, (2, 5000)).astype(np.float32)
    elif case == "clusters":
        c = rng.uniform(-80, 80, (2, 12, 1, 3))
        x = (c + rng.standard_normal((2, 12, 600, 3)) * 0.05).reshape(2, 7200, 3).astype(np.float32)
        x = x[:, rng.permutation(7200)]
    elif case == "select_all":
        x = _cloud(14, 1, 2500)
        npoint = 2500
    elif case == "more_than_all":
        x = _cloud(15, 2, 3000)
        npoint = 3100         # every minimum is 0 after 3000 samples: index 0 again and again, as the plain scan does
    elif case == "lidar":
        from ssf_slam_b200 import synth
        x = np.stack([it["pos1"] for it in synth.make_sequence(77, 2, 8192)]).astype(np.float32)
        npoint = 2048
    else:
        x = (rng.standard_normal((2, 8192, 3)) * 1e-3).astype(np.float32)
        x[:, 100] += 1e4      # one far outlier squeezes everything else into one cell of the curve
    got = pu.furthest_point_sample(_cuda(x), npoint).cpu().numpy()
    assert np.array_equal(got, po.c_fps(np.ascontiguousarray(x), npoint))


@pytest.mark.parametrize("Nq,Nr,k", [(2048, 8192, 16), (8192, 8192, 16), (8192, 2048, 7), (512, 256, 5), (128, 256, 8),
                                     (8192, 2048, 3), (1000, 1501, 16), (77, 33, 32), (3, 1, 1), (4096, 5000, 16)])
def test_knn(pu, Nq, Nr, k):
    B = 2
    ref = _cloud(Nq + Nr, B, Nr)
    q = _cloud(7 * Nq + 1, B, Nq, dup=False)
    q[:, : min(Nq, Nr) // 2] = ref[:, : min(Nq, Nr) // 2]  # queries coinciding with reference points (d = 0 ties)
    d, i = pu.knn(k, _cuda(q), _cuda(ref))
    od, oi = po.c_knn(k, q, ref)
    assert np.array_equal(i.cpu().numpy(), oi)
    assert np.array_equal(d.cpu().numpy(), od)


def test_knn_offset_matches_materialised_sum(pu):
    ref = _cloud(1, 2, 4096)
    q = _cloud(2, 2, 2048, dup=False)
    off = (np.random.default_rng(3).standard_normal(q.shape) * 0.3).astype(np.float32)
    _, i1 = pu.knn(16, _cuda(q), _cuda(ref), offset=_cuda(off))
    _, oi = po.c_knn(16, q + off, ref)
    assert np.array_equal(i1.cpu().numpy(), oi)


def test_knn_all_ties(pu):
    ref = np.zeros((1, 300, 3), np.float32)
    d, i = pu.knn(16, _cuda(np.zeros((1, 5, 3), np.float32)), _cuda(ref))
    assert np.array_equal(i.cpu().numpy()[0], np.tile(np.arange(16), (5, 1))) and float(d.abs().max()) == 0.0


def test_knn_rejects_bad_k(pu):
    from ssf_slam_b200._native import SsfError
    with pytest.raises(SsfError):
        pu.knn(9, torch.zeros(1, 4, 3).cuda(), torch.zeros(1, 4, 3).cuda())  # k > Nr
    with pytest.raises(SsfError):
        pu.knn(64, torch.zeros(1, 4, 3).cuda(), torch.zeros(1, 100, 3).cuda())


@pytest.mark.parametrize("N,S,radius,nsample", [(8192, 2048, 0.5, 16), (8192, 2048, 4.0, 32), (1501, 333, 2.0, 16),
                                                (65536, 2048, 1.0, 16), (100, 10, 0.001, 8)])
def test_ball_query(pu, N, S, radius, nsample):
    xyz = _cloud(N + S, 2, N, scale=20.0)
    new = xyz[:, :: max(1, N // S)][:, :S].copy()
    if radius < 0.01:
        new = new + 7.0  # nobody in range
    oi, oc = po.c_ball_query(radius, nsample, xyz, new)
    for use_index in (False, True) if N >= 64 else (False,):    # the brute-force scan and the spatial index: same output
        bi, bc = pu.ball_query(radius, nsample, _cuda(xyz), _cuda(new), return_count=True, use_index=use_index)
        assert np.array_equal(bc.cpu().numpy(), oc)
        assert np.array_equal(bi.cpu().numpy(), oi)
        assert np.array_equal(pu.ball_query(radius, nsample, _cuda(xyz), _cuda(new), use_index=use_index).cpu().numpy(), oi)   # no counts


def test_ball_query_many_queries_thread_per_query_kernel(pu):
    """B * S above the warp-per-query threshold (32768): the thread-per-query kernel; and just below it: the warp kernel."""
    for B, S in ((20, 2048), (16, 2048)):
        xyz = _cloud(77, B, 1024, scale=10.0)
        new = np.concatenate([xyz, xyz[:, ::-1] + 0.3], axis=1).copy()[:, :S]
        bi, bc = pu.ball_query(1.5, 16, _cuda(xyz), _cuda(new), return_count=True)
        oi, oc = po.c_ball_query(1.5, 16, xyz, new)
        assert np.array_equal(bc.cpu().numpy(), oc) and np.array_equal(bi.cpu().numpy(), oi)


@pytest.mark.parametrize("C,N,M,S", [(3, 8192, 2048, 16), (96, 8192, 8192, 16), (64, 2048, 8192, 7), (5, 100, 33, 3), (1, 7, 1, 1)])
def test_grouping_and_gather(pu, C, N, M, S):
    rng = np.random.default_rng(C * N + M)
    feat = rng.standard_normal((2, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (2, M, S)).astype(np.int32)
    got = pu.grouping_operation(_cuda(feat), _cuda(idx)).cpu().numpy()
    assert np.array_equal(got, po.c_group(feat, idx))
    g1 = pu.gather_operation(_cuda(feat), _cuda(idx[:, :, 0].copy())).cpu().numpy()
    assert np.array_equal(g1, po.gather_operation(torch.from_numpy(feat), torch.from_numpy(idx[:, :, 0].copy())).numpy())


@pytest.mark.parametrize("C,M,N", [(33, 500, 1000), (6, 8192, 8192), (5, 500, 1002), (3, 64, 40), (4, 2048, 8192)])
def test_three_interpolate(pu, C, M, N):
    """staged + four-points-per-thread path (N % 4 == 0), staged scalar path, plain path (N < M)"""
    rng = np.random.default_rng(9)
    feat = rng.standard_normal((2, C, M)).astype(np.float32)
    idx = rng.integers(0, M, (2, N, 3)).astype(np.int32)
    w = rng.random((2, N, 3)).astype(np.float32)
    got = pu.three_interpolate(_cuda(feat), _cuda(idx), _cuda(w)).cpu().numpy()
    want = po.three_interpolate(torch.from_numpy(feat), torch.from_numpy(idx), torch.from_numpy(w)).numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("L,C,n", [(4096 * 16, 1, 4096), (2048 * 16, 64, 2048), (1000, 5, 37), (50, 3, 200)])
def test_scatter(L, C, n):
    from ssf_slam_b200 import scatter
    rng = np.random.default_rng(L + C)
    src = rng.standard_normal((2, L, C)).astype(np.float32)
    index = rng.integers(0, n, (2, L)).astype(np.int64)
    index[:, 0] = n - 1
    sm = scatter.scatter_softmax(_cuda(src), _cuda(index), dim=1).cpu()
    ss = scatter.scatter_sum(_cuda(src), _cuda(index), dim=1).cpu()
    osm = po.scatter_softmax(torch.from_numpy(src), torch.from_numpy(index), dim=1)
    oss = po.scatter_sum(torch.from_numpy(src), torch.from_numpy(index), dim=1)
    assert ss.shape == oss.shape
    assert float((sm - osm).abs().max()) <= 2e-6   # fp32 exp/ordering tolerance
    assert float((ss - oss).abs().max()) <= 1e-4 * max(1.0, float(oss.abs().max()))
    # run-to-run determinism (no atomics in the accumulation)
    ss2 = scatter.scatter_sum(_cuda(src), _cuda(index), dim=1).cpu()
    assert torch.equal(ss, ss2)


@pytest.mark.parametrize("B,Nq,Nr,k,dup,off", [(2, 700, 1000, 16, False, False), (2, 2048, 8192, 16, True, True), (1, 8192, 8192, 7, True, False),
                                               (3, 513, 2048, 3, False, True), (1, 300, 16384, 5, False, False), (2, 64, 33, 16, False, False),
                                               (1, 4096, 16385, 16, True, False), (2, 3000, 40000, 16, True, True), (1, 2048, 65536, 7, False, False),
                                               (1, 1024, 100001, 16, True, False), (1, 512, 131072, 3, False, True)])
def test_knn_blocks_equals_brute_force(B, Nq, Nr, k, dup, off):
    """The block search returns exactly the brute-force scan's indices and distances (ties included)."""
    import torch
    from ssf_slam_b200 import _native as nat
    g = torch.Generator().manual_seed(Nq + Nr + k)
    ref = torch.randn(B, Nr, 3, generator=g) * torch.tensor([30.0, 20.0, 2.0])
    if dup:  # duplicated points: exact distance ties, the index order is observable
        ref[:, Nr // 2:] = ref[:, :Nr - Nr // 2].clone()
    query = ref[:, torch.randperm(Nr, generator=g)[:Nq] % Nr].clone() if Nq <= Nr else torch.randn(B, Nq, 3, generator=g) * 20
    query[:, ::3] += torch.randn(B, (Nq + 2) // 3, 3, generator=g)
    qadd = (torch.randn(B, Nq, 3, generator=g) * 0.5).cuda() if off else None
    ref, query = ref.cuda().contiguous(), query.cuda().contiguous()
    L = nat.lib()
    d0 = torch.empty(B, Nq, k, device="cuda"); i0 = torch.empty(B, Nq, k, dtype=torch.int32, device="cuda")
    d1 = torch.full((B, Nq, k), -1.0, device="cuda"); i1 = torch.full((B, Nq, k), -1, dtype=torch.int32, device="cuda")
    nat.check(L.ssf_knn_offset(k, nat.ptr(query), nat.ptr(qadd), nat.ptr(ref), B, Nq, Nr, nat.ptr(d0), nat.ptr(i0), nat.stream()))
    ws = torch.empty(int(L.ssf_knn_blocks_workspace_floats(B, Nr)), device="cuda")
    nat.check(L.ssf_knn_blocks_build(nat.ptr(ref), B, Nr, nat.ptr(ws), nat.stream()))
    nat.check(L.ssf_knn_blocks_search(k, nat.ptr(query), nat.ptr(qadd), nat.ptr(ws), B, Nq, Nr, nat.ptr(d1), nat.ptr(i1), nat.stream()))
    torch.cuda.synchronize()
    assert torch.equal(i0, i1)
    assert torch.equal(d0, d1)


@pytest.mark.parametrize("B,L,C,n_seg,hot", [(2, 4096, 64, 300, False), (1, 3000, 128, 50, True), (2, 1000, 5, 400, False), (1, 64, 256, 7, True),
                                             (1, 40000, 8, 10, True)])
def test_segment_softmax_sum_fused(B, L, C, n_seg, hot):
    """Backward-cost kernel (softmax over each segment's logits, weighted sum of its rows) vs the dense definition;
    `hot` concentrates rows on a few keys so that segments exceed one warp batch (32 rows)."""
    import torch
    from ssf_slam_b200 import functional as F_
    g = torch.Generator().manual_seed(L + C)
    key = torch.randint(0, 3 if hot else n_seg, (B, L), generator=g, dtype=torch.int32)
    logit = torch.randn(B, L, generator=g) * 3
    val = torch.randn(B, L, C, generator=g)
    csr = F_.build_csr(key.cuda(), n_seg)
    got = F_.segment_softmax_sum(logit.cuda(), val.cuda().contiguous(), csr, n_seg).cpu().double()
    want = torch.zeros(B, n_seg, C, dtype=torch.float64)
    for b in range(B):
        for j in key[b].unique().tolist():
            m = key[b] == j
            w = torch.softmax(logit[b][m].double(), dim=0)
            want[b, j] = (w[:, None] * val[b][m].double()).sum(0)
    assert float((got - want).abs().max()) < 1e-5
    again = F_.segment_softmax_sum(logit.cuda(), val.cuda().contiguous(), F_.build_csr(key.cuda(), n_seg), n_seg).cpu().double()
    assert torch.equal(got, again)   # deterministic


@pytest.mark.parametrize("B,Nq,Nr,k", [(2, 300, 20000, 16), (1, 2048, 65536, 16), (2, 77, 17000, 5), (1, 33, 16385, 32)])
def test_knn_warp_scan_large_cloud_few_queries(pu, B, Nq, Nr, k):
    """Nr above the block-index limit with few queries: the warp-per-two-queries scan; same (distance, index) order as the oracle,
    with exact duplicates in the cloud and queries on reference points."""
    ref = _cloud(Nq + Nr, B, Nr)
    q = _cloud(3 * Nq + 1, B, Nq, dup=False)
    q[:, : Nq // 2] = ref[:, : Nq // 2]
    d, i = pu.knn(k, _cuda(q), _cuda(ref))
    od, oi = po.c_knn(k, q, ref)
    assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od)
    off = (np.random.default_rng(4).standard_normal(q.shape) * 0.3).astype(np.float32)
    _, i2 = pu.knn(k, _cuda(q), _cuda(ref), offset=_cuda(off))
    assert np.array_equal(i2.cpu().numpy(), po.c_knn(k, q + off, ref)[1])


def test_config5_dense_stress_n65536(pu):
    """BASELINE config 5 (N = 65536 per cloud): the full 65536 x 65536 k = 16 search (every query; the oracle checks a strided
    query subsample and all rows that sit on duplicated points), FPS 65536 -> 2048, and the ball-query radius sweep
    0.5 / 1 / 2 / 4 m around the FPS centres, against the plain-C oracle; lidar-like anisotropic cloud with exact duplicates."""
    N = 65536
    rng = np.random.default_rng(65536)
    xyz = (rng.standard_normal((1, N, 3)) * np.array([30.0, 20.0, 2.0])).astype(np.float32)
    xyz[:, N // 2:N // 2 + 2048] = xyz[:, :2048]          # exact duplicates -> exact ties
    x = _cuda(xyz)
    d, i = pu.knn(16, x, x)
    rows = np.unique(np.concatenate([np.arange(0, N, 97), np.arange(0, 256), np.arange(N // 2, N // 2 + 256)]))
    od, oi = po.c_knn(16, xyz[:, rows], xyz)
    assert np.array_equal(i.cpu().numpy()[:, rows], oi) and np.array_equal(d.cpu().numpy()[:, rows], od)
    fps = pu.furthest_point_sample(x, 2048)
    assert np.array_equal(fps.cpu().numpy(), po.c_fps(xyz, 2048))
    cent = np.ascontiguousarray(xyz[0][fps.cpu().numpy()[0].astype(np.int64)][None])
    index = pu.build_index(x)                      # one build shared by the whole sweep (and by a kNN)
    d2, i2 = pu.knn(16, _cuda(cent), x, index=index)
    od2, oi2 = po.c_knn(16, cent, xyz)
    assert np.array_equal(i2.cpu().numpy(), oi2) and np.array_equal(d2.cpu().numpy(), od2)
    for r in (0.5, 1.0, 2.0, 4.0):
        oi, oc = po.c_ball_query(r, 16, xyz, cent)
        for kw in (dict(index=index), dict(use_index=False)):
            bi, bc = pu.ball_query(r, 16, x, _cuda(cent), return_count=True, **kw)
            assert np.array_equal(bc.cpu().numpy(), oc) and np.array_equal(bi.cpu().numpy(), oi), (r, kw)
    oi, oc = po.c_ball_query(30.0, 32, xyz, cent[:, :64])      # a ball that holds most of the cloud
    bi, bc = pu.ball_query(30.0, 32, x, _cuda(cent[:, :64]), return_count=True, index=index)
    assert np.array_equal(bc.cpu().numpy(), oc) and np.array_equal(bi.cpu().numpy(), oi)


def test_error_behaviour_python_exceptions(pu):
    """Errors are Python exceptions (the reference's extension raises too; there are no status codes at the B-op boundary,
    SURVEY 8(b)): empty inputs, CPU tensors (no CPU fallback), arguments out of range."""
    from ssf_slam_b200._native import SsfError
    z = lambda *s: torch.zeros(*s, device="cuda")
    zi = lambda *s: torch.zeros(*s, dtype=torch.int32, device="cuda")
    with pytest.raises(SsfError):
        pu.furthest_point_sample(z(1, 0, 3), 4)                       # empty cloud
    with pytest.raises(SsfError):
        pu.knn(3, z(1, 0, 3), z(1, 10, 3))                            # no queries
    with pytest.raises(SsfError):
        pu.knn(33, z(1, 4, 3), z(1, 100, 3))                          # k > 32
    with pytest.raises(SsfError):
        pu.ball_query(1.0, 0, z(1, 10, 3), z(1, 2, 3))                # nsample = 0
    with pytest.raises(SsfError):
        pu.grouping_operation(z(1, 0, 8), zi(1, 2, 2))                # no channels
    with pytest.raises(SsfError):
        pu.furthest_point_sample(torch.zeros(1, 16, 3), 4)            # CPU tensor: the product path has no CPU fallback
    with pytest.raises(SsfError):
        pu.knn(3, torch.zeros(1, 4, 3), torch.zeros(1, 10, 3))
    # ragged / tiny but valid inputs still work
    assert pu.furthest_point_sample(z(2, 1, 3), 1).tolist() == [[0], [0]]
    d, i = pu.knn(1, z(1, 1, 3), z(1, 1, 3))
    assert i.tolist() == [[[0]]] and float(d.abs().max()) == 0.0
    assert pu.ball_query(1.0, 3, z(1, 1, 3), z(1, 1, 3)).tolist() == [[[0, 0, 0]]]


def test_query_and_group_modules(pu):
    """`QueryAndGroup` / `GroupAll` (constructed by the reference at ASF/utils/soflow.py:1520-1523): ball query + grouping with
    centre-relative coordinates, against the same composition of the oracle's operators."""
    rng = np.random.default_rng(8)
    xyz = rng.uniform(-10, 10, (2, 3000, 3)).astype(np.float32)
    new = xyz[:, ::7][:, :400].copy()
    feat = rng.standard_normal((2, 6, 3000)).astype(np.float32)
    got = pu.QueryAndGroup(1.5, 16)(_cuda(xyz), _cuda(new), _cuda(feat)).cpu().numpy()
    oi, _ = po.c_ball_query(1.5, 16, xyz, new)
    g_xyz = po.c_group(np.ascontiguousarray(xyz.transpose(0, 2, 1)), oi) - new.transpose(0, 2, 1)[..., None]
    want = np.concatenate([g_xyz, po.c_group(feat, oi)], axis=1)
    assert got.shape == (2, 9, 400, 16) and np.array_equal(got, want)
    assert np.array_equal(pu.QueryAndGroup(1.5, 16, use_xyz=False)(_cuda(xyz), _cuda(new), _cuda(feat)).cpu().numpy(), want[:, 3:])
    ga = pu.GroupAll()(_cuda(xyz), _cuda(new), _cuda(feat)).cpu().numpy()
    assert ga.shape == (2, 9, 1, 3000) and np.array_equal(ga[:, :3, 0], xyz.transpose(0, 2, 1)) and np.array_equal(ga[:, 3:, 0], feat)


def test_index_paths_random_sizes(pu):
    """Seeded sweep over ragged sizes around every dispatch threshold (scan / one-level / two-level index, warp scan, block
    counts that are not multiples of 32, super-block counts that are not multiples of 32): kNN and ball query against the C
    oracle, with duplicates in every cloud."""
    rng = np.random.default_rng(4242)
    sizes = [65, 255, 256, 257, 511, 512, 513, 1000, 2047, 2049, 4097, 16384, 16385, 16417, 20011, 32769, 33333, 65537, 100003]
    for Nr in sizes:
        B = 1 if Nr > 20000 else 2
        Nq = int(rng.integers(33, 700))
        k = int(rng.choice([1, 3, 7, 16, 32]))
        ref = (rng.standard_normal((B, Nr, 3)) * np.array([30.0, 20.0, 2.0])).astype(np.float32)
        ref[:, Nr // 3:Nr // 3 + 40] = ref[:, :40]
        q = np.concatenate([ref[:, :Nq // 2], (rng.standard_normal((B, Nq - Nq // 2, 3)) * 25).astype(np.float32)], axis=1)
        index = pu.build_index(_cuda(ref))
        d, i = pu.knn(k, _cuda(q), _cuda(ref), index=index)
        od, oi = po.c_knn(k, q, ref)
        assert np.array_equal(i.cpu().numpy(), oi) and np.array_equal(d.cpu().numpy(), od), ("knn", Nr, Nq, k)
        d2, i2 = pu.knn(k, _cuda(q), _cuda(ref))          # the default dispatch for this shape
        assert np.array_equal(i2.cpu().numpy(), oi), ("knn default", Nr, Nq, k)
        r = float(rng.choice([0.3, 1.0, 3.0]))
        ns = int(rng.choice([1, 8, 16, 32]))
        bi, bc = pu.ball_query(r, ns, _cuda(ref), _cuda(q), return_count=True, index=index)
        wi, wc = po.c_ball_query(r, ns, ref, q)
        assert np.array_equal(bc.cpu().numpy(), wc) and np.array_equal(bi.cpu().numpy(), wi), ("ball", Nr, Nq, r, ns)
