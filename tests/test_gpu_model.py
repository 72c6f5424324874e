"""GPU parity: fused layers (teacher-forced against the oracle port's intermediates) and the whole TFlow forward
against the committed reference goldens.  Tolerances: FPS / kNN indices exact; flow max-abs <= 1e-4 m (north star)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import point_ops as po  # noqa: E402  (checker only)
from oracle import tflow_port as tp  # noqa: E402  (checker only)

FLOW_TOL = 1e-4


def _pm(x):  # [B,C,N] cpu -> [B,N,C] cuda
    return x.permute(0, 2, 1).contiguous().cuda()


def _cm(x):  # [B,N,C] cuda -> [B,C,N] cpu
    return x.permute(0, 2, 1).contiguous().cpu()


@pytest.fixture(scope="module")
def setup(golden_dir, oracle_c):
    from ssf_slam_b200.model import TFlow, prepare_weights
    g = np.load(os.path.join(golden_dir, "tflow_n2048.npz"))
    sd = tp.random_init_state_dict(int(g["weight_seed"]))
    pc1 = torch.from_numpy(g["pos1"].T.copy()).unsqueeze(0)
    pc2 = torch.from_numpy(g["pos2"].T.copy()).unsqueeze(0)
    (flows, fps), inter = tp.tflow_forward(sd, pc1, pc2, return_intermediates=True)
    net = TFlow()
    net.load_state_dict(sd, strict=True)
    W = prepare_weights(sd, torch.device("cuda:0"))
    return dict(g=g, sd=sd, pc1=pc1, pc2=pc2, flows=flows, fps=fps, inter=inter, net=net, W=W)


def test_linear_matches_conv1d(setup):
    from ssf_slam_b200 import functional as F_
    sd, W = setup["sd"], setup["W"]
    x = setup["pc1"]
    want = tp.leaky_conv1d(sd, "point_conv.1", tp.leaky_conv1d(sd, "point_conv.0", x))
    got = F_.linear(F_.linear(_pm(x), W["pc0"], 32, act=2), W["pc1"], 32, act=2)
    assert float((_cm(got) - want).abs().max()) < 1e-5
    assert float((_cm(got) - setup["inter"]["f1"][0]).abs().max()) < 1e-5


@pytest.mark.parametrize("lvl", [1, 2, 3, 4])
def test_set_abstraction_teacher_forced(setup, lvl):
    from ssf_slam_b200.model import set_abstraction_pm
    name, npoint, nsample = tp.SA_CFG[lvl - 1]
    inter = setup["inter"]
    xyz, feats = inter["pcs1"][lvl - 1], inter["f1"][lvl - 1]
    nx, nf, fi = set_abstraction_pm(setup["W"][name], npoint, nsample, _pm(xyz), _pm(feats))
    assert torch.equal(fi.cpu(), inter["fps"][lvl - 1])
    assert torch.equal(_cm(nx), inter["pcs1"][lvl])
    err = float((_cm(nf) - inter["f1"][lvl]).abs().max())
    assert err < 2e-5 * max(1.0, float(inter["f1"][lvl].abs().max())), err


@pytest.fixture(scope="module")
def setup8k(golden_dir, oracle_c):
    """The N = 8192 golden pair (BASELINE configs 1/2/4 size) with every intermediate of the oracle port, so that each fused
    layer -- including the tcgen05 m = 64 cost volume of levels 1 and 0, the dominant kernel -- is fed exactly what the
    reference fed its own layer (identical neighbour sets) and compared output by output."""
    from ssf_slam_b200.model import prepare_weights
    g = np.load(os.path.join(golden_dir, "tflow_n8192.npz"))
    sd = tp.random_init_state_dict(int(g["weight_seed"]))
    pc1 = torch.from_numpy(g["pos1"].T.copy()).unsqueeze(0)
    pc2 = torch.from_numpy(g["pos2"].T.copy()).unsqueeze(0)
    (flows, fps), inter = tp.tflow_forward(sd, pc1, pc2, return_intermediates=True)
    # the port on THIS host's CPU against the golden the unmodified reference wrote in the build container: bit-identical there
    # (oracle/gen_golden.py asserts it), last-ulp differences of the CPU's conv kernels elsewhere
    for i in range(4):
        assert float(np.abs(flows[i][0].numpy() - g["flow%d" % i]).max()) <= 1e-5
    for i in range(3):
        assert np.array_equal(fps[i][0].numpy(), g["fps%d" % (i + 1)])
    return dict(sd=sd, inter=inter, W=prepare_weights(sd, torch.device("cuda:0")))


@pytest.mark.parametrize("lvl", [2, 1, 0])
def test_set_upconv_teacher_forced_n8192(setup8k, lvl):
    """a5: su2 / su1 / su0 (both clouds) against the oracle port's own outputs at N = 8192."""
    from ssf_slam_b200.model import set_upconv_pm
    inter = setup8k["inter"]
    io = inter["in%d" % lvl]
    for pcs, f, prev, want in ((inter["pcs1"], inter["f1"], io["prev1"], io["u1"]), (inter["pcs2"], inter["f2"], io["prev2"], io["u2"])):
        got = set_upconv_pm(setup8k["W"][io["su"]], 16, _pm(pcs[lvl]), _pm(pcs[lvl + 1]), _pm(f[lvl]), _pm(prev))
        err = float((_cm(got) - want).abs().max())
        assert err < 2e-5 * max(1.0, float(want.abs().max())), (lvl, err)


@pytest.mark.parametrize("lvl", [2, 1, 0])
def test_cost_volume_teacher_forced_n8192(setup8k, lvl):
    """a6: flow2_r (un-fused tcgen05 path, m = 128) and flow1_r / flow0_r (the fused tcgen05 kernel `cost_volume_tc64`, m = 64)
    against the oracle port at N = 8192: warp, both kNNs, both branches, attention, weightnet, forward / backward cost, flow head."""
    from ssf_slam_b200.model import cost_volume_pm, point_warping_pm
    inter = setup8k["inter"]
    io = inter["in%d" % lvl]
    p1, p2 = inter["pcs1"][lvl], inter["pcs2"][lvl]
    want_warp = tp.point_warping(p1, p2, io["coarse"], io["k_warp"])
    warped = point_warping_pm(_pm(p1), _pm(p2), _pm(io["coarse"]), io["k_warp"])
    assert float((_cm(warped) - want_warp).abs().max()) < 1e-5
    got = cost_volume_pm(setup8k["W"][io["fr"]], _pm(p1), _pm(p2), _pm(want_warp), _pm(io["u1"]), _pm(io["cfu"]), _pm(io["u2"]),
                         _pm(io["cbu"]), sf=_pm(io["coarse"]), sf_feat=_pm(io["sf_feat"]))
    want = inter["l%d" % lvl]
    for n, w_, g_ in zip(("cost_fwd", "cost_bwd", "feats", "flow"), want, got):
        w_ = w_ if n != "cost_bwd" else torch.nn.functional.pad(w_, (0, g_.shape[1] - w_.shape[2]))
        err = float((_cm(g_) - w_).abs().max())
        assert err < 3e-5 * max(1.0, float(w_.abs().max())), (lvl, n, err)


def test_cost_volume_tc64_kernel_vs_oracle_internals(setup8k):
    """The fused tcgen05 kernel's raw outputs at level 0 (N1 = 8192, m = 64) against the oracle port's internals: the forward
    cost, and the warped-branch logits gw and rows Cw that feed the segmented softmax (soflow.py:397-469)."""
    from ssf_slam_b200 import functional as F_
    inter, sd, W = setup8k["inter"], setup8k["sd"], setup8k["W"]["flow0_r"]
    io = inter["in0"]
    p1, p2 = inter["pcs1"][0], inter["pcs2"][0]
    warp = tp.point_warping(p1, p2, io["coarse"], io["k_warp"])
    _, aux = tp.cost_volume(sd, "flow0_r.cost", 16, True, p1, p2, warp, torch.cat([io["u1"], io["cfu"]], 1),
                            torch.cat([io["u2"], io["cbu"]], 1), io["coarse"], io["sf_feat"], return_aux=True)
    m = 64
    idx, idxw = aux["idx"].int().cuda().contiguous(), aux["idxw"].int().cuda().contiguous()
    Hab = F_.dense_tc(W["Hab_img"], 2 * m, W["D"], x1=_pm(io["u1"]), x2=_pm(io["cfu"]), bias=W["bab"])
    Gab = F_.dense_tc(W["Gab_img"], 2 * m, W["D"], x1=_pm(io["u2"]), x2=_pm(io["cbu"]))
    H3 = F_.dense_tc(W["H3_img"], m, W["F"], x1=_pm(io["sf_feat"]), bias=W["b3"])
    cost_fwd, cost_fwd_cm, gw, Cw = F_.cost_volume(Gab, Hab, W, H3, _pm(p1), _pm(p2), idx, idxw, m)
    want_fwd = inter["l0"][0]                                                     # [1, m, N1]
    assert float((cost_fwd_cm.cpu() - want_fwd).abs().max()) < 3e-5 * max(1.0, float(want_fwd.abs().max()))
    assert torch.equal(cost_fwd.permute(0, 2, 1).contiguous(), cost_fwd_cm)
    want_gw = aux["gw"].permute(0, 3, 2, 1).reshape(1, -1)                        # [1, N1*16]
    want_cw = aux["cw"].permute(0, 3, 2, 1).reshape(1, -1, m)
    assert float((gw.cpu() - want_gw).abs().max()) < 3e-5 * max(1.0, float(want_gw.abs().max()))
    assert float((Cw.cpu() - want_cw).abs().max()) < 3e-5 * max(1.0, float(want_cw.abs().max()))


class _CudaOps:
    """`ops` namespace for oracle/tflow_port.py whose point operators are the CUDA drop-ins behind the B-op boundary
    (ssf_slam_b200.pointnet2_utils / scatter, called with the reference's argument layouts); CPU tensors in and out."""

    def __init__(self):
        from ssf_slam_b200 import pointnet2_utils as pu, scatter
        self.pu, self.sc = pu, scatter

    def _c(self, t):
        return t.contiguous().cuda()

    def furthest_point_sample(self, xyz, n):
        return self.pu.furthest_point_sample(self._c(xyz), n).cpu()

    def gather_operation(self, f, idx):
        return self.pu.gather_operation(self._c(f), self._c(idx.int())).cpu()

    def knn(self, k, q, r):
        d, i = self.pu.knn(k, self._c(q), self._c(r))
        return d.cpu(), i.cpu()

    def three_nn(self, q, r):
        d, i = self.pu.three_nn(self._c(q), self._c(r))
        return d.cpu(), i.cpu()

    def grouping_operation(self, f, idx):
        return self.pu.grouping_operation(self._c(f.float()), self._c(idx.int())).cpu()

    def scatter_softmax(self, src, index, dim=1):
        return self.sc.scatter_softmax(self._c(src), self._c(index), dim=dim).cpu()

    def scatter_sum(self, src, index, dim=1, dim_size=None):
        return self.sc.scatter_sum(self._c(src), self._c(index), dim=dim, dim_size=dim_size).cpu()


def test_reference_side_code_through_the_bop_boundary(setup, monkeypatch):
    """B-op: the reference-side model code (the oracle port = the reference's torch layers, pinned bit-exact to the unmodified
    reference) run with the CUDA drop-ins as its `lib.pointnet2_utils` / `torch_scatter` -- the integration INTEGRATION.md
    describes, on the GPU box.  FPS indices equal the reference golden; flows within the north-star tolerance (only the scatter's summation order and expf differ)."""
    g = setup["g"]
    monkeypatch.setattr(tp, "ops", _CudaOps())
    flows, fps = tp.tflow_forward(setup["sd"], setup["pc1"], setup["pc2"])
    for i in range(3):
        assert np.array_equal(fps[i][0].numpy(), g["fps%d" % (i + 1)])
    errs = [float(np.abs(flows[i][0].numpy() - g["flow%d" % i]).max()) for i in range(4)]
    print("port over CUDA B-ops vs reference golden, flow max-abs error per level:", errs)
    assert max(errs) <= FLOW_TOL, errs


@pytest.mark.parametrize("lvl,name", [(3, "su3")])
def test_set_upconv_teacher_forced(setup, lvl, name):
    from ssf_slam_b200.model import set_upconv_pm
    inter, sd = setup["inter"], setup["sd"]
    want = tp.set_upconv(sd, name, 16, inter["pcs1"][lvl], inter["pcs1"][lvl + 1], inter["f1"][lvl], inter["f1"][lvl + 1])
    got = set_upconv_pm(setup["W"][name], 16, _pm(inter["pcs1"][lvl]), _pm(inter["pcs1"][lvl + 1]), _pm(inter["f1"][lvl]),
                        _pm(inter["f1"][lvl + 1]))
    err = float((_cm(got) - want).abs().max())
    assert err < 2e-5 * max(1.0, float(want.abs().max())), err


def test_upsample_and_warp_teacher_forced(setup):
    from ssf_slam_b200.model import point_warping_pm, upsample_pm
    inter = setup["inter"]
    cf, cb, ff, flow = inter["l3"]
    for val, k in ((flow, 5), (ff, 5), (cf, 3)):
        want = tp.upsample_flow(inter["pcs1"][2], inter["pcs1"][3], val, k=k)
        got = upsample_pm(_pm(inter["pcs1"][2]), _pm(inter["pcs1"][3]), _pm(val), k)
        assert float((_cm(got) - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))
    coarse = tp.upsample_flow(inter["pcs1"][2], inter["pcs1"][3], flow, k=5)
    want = tp.point_warping(inter["pcs1"][2], inter["pcs2"][2], coarse, 5)
    got = point_warping_pm(_pm(inter["pcs1"][2]), _pm(inter["pcs2"][2]), _pm(coarse), 5)
    assert float((_cm(got) - want).abs().max()) < 1e-5


def test_cost_volume_level3_teacher_forced(setup):
    """flow3_r: no flow / no sf_feat inputs.  Inputs come from the oracle so neighbour sets are identical."""
    from ssf_slam_b200.model import cost_volume_pm
    inter, sd = setup["inter"], setup["sd"]
    u1 = tp.set_upconv(sd, "su3", 16, inter["pcs1"][3], inter["pcs1"][4], inter["f1"][3], inter["f1"][4])
    u2 = tp.set_upconv(sd, "su3", 16, inter["pcs2"][3], inter["pcs2"][4], inter["f2"][3], inter["f2"][4])
    want = tp.cost_volume(sd, "flow3_r.cost", 16, False, inter["pcs1"][3], inter["pcs2"][3], None, u1, u2)
    got = cost_volume_pm(setup["W"]["flow3_r"], _pm(inter["pcs1"][3]), _pm(inter["pcs2"][3]), None, _pm(u1), None, _pm(u2), None)
    names = ("cost_fwd", "cost_bwd", "feats", "flow")
    for n, w_, g_ in zip(names, want, got):
        w_ = w_ if n != "cost_bwd" else torch.nn.functional.pad(w_, (0, g_.shape[1] - w_.shape[2]))
        err = float((_cm(g_) - w_).abs().max())
        assert err < 3e-5 * max(1.0, float(w_.abs().max())), (n, err)


def test_cost_volume_level2_teacher_forced(setup):
    """flow2_r: warping, flow-shifted kNN, sf_feat and the 2-segment feature inputs."""
    from ssf_slam_b200.model import cost_volume_pm, point_warping_pm
    inter, sd = setup["inter"], setup["sd"]
    p1, p2, p1s = inter["pcs1"][2], inter["pcs2"][2], inter["pcs1"][3]
    u31 = tp.set_upconv(sd, "su3", 16, inter["pcs1"][3], inter["pcs1"][4], inter["f1"][3], inter["f1"][4])
    u32 = tp.set_upconv(sd, "su3", 16, inter["pcs2"][3], inter["pcs2"][4], inter["f2"][3], inter["f2"][4])
    u1 = tp.set_upconv(sd, "su2", 16, p1, p1s, inter["f1"][2], u31)
    u2 = tp.set_upconv(sd, "su2", 16, p2, inter["pcs2"][3], inter["f2"][2], u32)
    cf, cb, ff, flow = inter["l3"]
    coarse = tp.upsample_flow(p1, p1s, flow, k=5)
    sf_feat = tp.upsample_flow(p1, p1s, ff, k=5)
    cfu = tp.leaky_conv1d(sd, "deconv3_2", tp.upsample_flow(p1, p1s, cf))
    cbu = tp.leaky_conv1d(sd, "deconv3_2", tp.upsample_flow(p1, p1s, cb))
    want = tp.refine_flow(sd, "flow2_r", 16, True, p1, p2, torch.cat([u1, cfu], 1), torch.cat([u2, cbu], 1), 5, coarse, sf_feat)
    warped = point_warping_pm(_pm(p1), _pm(p2), _pm(coarse), 5)
    got = cost_volume_pm(setup["W"]["flow2_r"], _pm(p1), _pm(p2), warped, _pm(u1), _pm(cfu), _pm(u2), _pm(cbu),
                         sf=_pm(coarse), sf_feat=_pm(sf_feat))
    for n, w_, g_ in zip(("cost_fwd", "cost_bwd", "feats", "flow"), want, got):
        w_ = w_ if n != "cost_bwd" else torch.nn.functional.pad(w_, (0, g_.shape[1] - w_.shape[2]))
        err = float((_cm(g_) - w_).abs().max())
        assert err < 3e-5 * max(1.0, float(w_.abs().max())), (n, err)
    for n, w_ in zip(("cost_fwd", "cost_bwd", "feats", "flow"), want):
        assert float((inter["l2"][("cost_fwd", "cost_bwd", "feats", "flow").index(n)] - w_).abs().max()) == 0.0


@pytest.mark.parametrize("n", [2048, 8192])
def test_tflow_end_to_end_vs_reference_golden(setup, golden_dir, n):
    g = np.load(os.path.join(golden_dir, "tflow_n%d.npz" % n))
    net = setup["net"]
    pc1 = torch.from_numpy(g["pos1"].T.copy()).unsqueeze(0).cuda()
    pc2 = torch.from_numpy(g["pos2"].T.copy()).unsqueeze(0).cuda()
    flows, fps = net(pc1, pc2)
    for i in range(3):
        assert np.array_equal(fps[i][0].cpu().numpy(), g["fps%d" % (i + 1)]), "fps level %d" % (i + 1)
    errs = [float(np.abs(flows[i][0].cpu().numpy() - g["flow%d" % i]).max()) for i in range(4)]
    print("flow max-abs error per level (fine -> coarse):", errs)
    assert flows[0].shape == (1, 3, n) and flows[1].shape == (1, 3, 2048)
    for e in errs:
        assert e <= FLOW_TOL, errs


def test_batch_independence(setup, golden_dir):
    """B=2 batch == two B=1 calls (the reference forward is batch-independent on CPU)."""
    g = np.load(os.path.join(golden_dir, "tflow_n2048.npz"))
    net = setup["net"]
    a = torch.from_numpy(g["pos1"].T.copy()).unsqueeze(0).cuda()
    b = torch.from_numpy(g["pos2"].T.copy()).unsqueeze(0).cuda()
    f1, _ = net(a, b)
    f2, _ = net(torch.cat([a, b]), torch.cat([b, a]))
    assert torch.equal(f1[0][0], f2[0][0])
    f3, _ = net(b, a)
    assert torch.equal(f3[0][0], f2[0][1])


def test_state_dict_with_dataparallel_prefix(setup):
    from ssf_slam_b200.model import prepare_weights
    sd = {"module." + k: v for k, v in setup["sd"].items()}
    W = prepare_weights(sd, torch.device("cuda:0"))
    assert torch.equal(W["pc0"], setup["W"]["pc0"])


def test_frontend_pipeline_host_buffers(setup, golden_dir):
    from oracle import frontend as ofe
    from ssf_slam_b200.frontend import SceneFlowFrontEnd
    g = np.load(os.path.join(golden_dir, "tflow_n2048.npz"))
    fe = SceneFlowFrontEnd(setup["net"], tau=0.10)
    out = fe.process(g["pos1"][None], g["pos2"][None], return_flow=True)
    flow = out["flow"][0].numpy()
    assert np.abs(flow.T - g["flow0"]).max() <= FLOW_TOL
    want = ofe.masker_spec(g["pos1"], flow, 0.10)  # mask/pose spec on the very flow the GPU produced
    assert np.array_equal(out["mask"][0].numpy(), want["mask"])
    assert np.array_equal(out["odom"][0].numpy(), want["odom"])


def test_tflow_n16384_vs_reference_golden_and_seg_masker(setup, golden_dir):
    """BASELINE config 3 (Seg pipeline, N = 16384): golden written by the unmodified reference TFlow at N = 16384
    (oracle/gen_golden.py): FPS exact, flows <= 1e-4; then the Seg masker (semantic seed + per-instance voting) and the pose on
    the GPU's own flow, bit-exact against the specification.  The SIMT realisation must agree with the tensor-core one too."""
    from oracle import frontend as ofe
    from ssf_slam_b200 import functional as F_, synth
    from ssf_slam_b200.frontend import SceneFlowFrontEnd
    g = np.load(os.path.join(golden_dir, "tflow_n16384.npz"))
    net = setup["net"]
    pc1 = torch.from_numpy(g["pos1"].T.copy()).unsqueeze(0).cuda()
    pc2 = torch.from_numpy(g["pos2"].T.copy()).unsqueeze(0).cuda()
    flows_tc, fps_tc = net(pc1, pc2)
    for i in range(3):
        assert np.array_equal(fps_tc[i][0].cpu().numpy(), g["fps%d" % (i + 1)]), "fps level %d" % (i + 1)
    errs = [float(np.abs(flows_tc[i][0].cpu().numpy() - g["flow%d" % i]).max()) for i in range(4)]
    print("N=16384 flow max-abs error per level vs the reference golden:", errs)
    assert flows_tc[0].shape == (1, 3, 16384) and max(errs) <= FLOW_TOL
    F_.USE_TC = False
    try:
        flows_simt, fps_simt = net(pc1, pc2)
    finally:
        F_.USE_TC = True
    for a, b in zip(fps_tc, fps_simt):
        assert torch.equal(a, b)
    assert max(float(np.abs(flows_simt[i][0].cpu().numpy() - g["flow%d" % i]).max()) for i in range(4)) <= FLOW_TOL
    fe = SceneFlowFrontEnd(net, tau=0.10, movable=synth.MOVABLE_CLASSES)
    n_inst = int(g["inst"].max()) + 1
    out = fe.process(g["pos1"][None], g["pos2"][None], sem=g["sem"][None], inst=g["inst"][None], n_inst=n_inst, return_flow=True)
    flow = out["flow"][0].numpy()
    want = ofe.masker_spec(g["pos1"], flow, 0.10, sem=g["sem"], inst=g["inst"], movable=synth.MOVABLE_CLASSES)
    assert np.array_equal(out["mask"][0].numpy(), want["mask"])
    assert np.array_equal(out["odom"][0].numpy(), want["odom"])


def test_frontend_cuda_graph_replay_equals_eager(setup, golden_dir):
    """The captured CUDA graph of the whole step gives the same masks / poses / flows as the eager launches, for
    several different inputs replayed through the same graph."""
    from ssf_slam_b200 import synth
    from ssf_slam_b200.frontend import SceneFlowFrontEnd
    eager = SceneFlowFrontEnd(setup["net"], tau=0.10)
    graph = SceneFlowFrontEnd(setup["net"], tau=0.10, use_graph=True)
    for seed in (11, 12, 13):
        it = synth.make_pair(seed, 2048)
        a = {k: v.clone() for k, v in eager.process(it["pos1"][None], it["pos2"][None], return_flow=True).items()}
        b = graph.process(it["pos1"][None], it["pos2"][None], return_flow=True)
        assert torch.equal(a["mask"], b["mask"]) and torch.equal(a["odom"], b["odom"]) and torch.equal(a["flow"], b["flow"])


def _sub_sd(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def test_layer_classes_are_dropins(setup):
    """B-layer boundary: the layer classes constructed and called exactly as ASF/TFlowV3_Occlussion.py:70-97,113-187 does
    (reference [B,C,N] layouts, own state_dict), checked against the oracle's restatement of the reference layers."""
    from ssf_slam_b200 import model as M
    inter, sd = setup["inter"], setup["sd"]
    cu = lambda t: t.cuda().contiguous()
    # PointNetSetAbstraction (sa2)
    sa = M.PointNetSetAbstraction(npoint=512, radius=2.0, nsample=16, in_channel=64, mlp=[64, 64, 128], group_all=False)
    sa.load_state_dict(_sub_sd(sd, "sa2."), strict=True)
    nx, nf, fi = sa(cu(inter["pcs1"][1]), cu(inter["f1"][1]))
    assert torch.equal(fi.cpu(), inter["fps"][1]) and torch.equal(nx.cpu(), inter["pcs1"][2])
    assert float((nf.cpu() - inter["f1"][2]).abs().max()) < 2e-5 * max(1.0, float(inter["f1"][2].abs().max()))
    # PointNetSetUpConv (su3)
    su = M.PointNetSetUpConv(nsample=16, radius=2.4, f1_channel=256, f2_channel=512, mlp=[256, 256], mlp2=[256, 256])
    su.load_state_dict(_sub_sd(sd, "su3."), strict=True)
    want = tp.set_upconv(sd, "su3", 16, inter["pcs1"][3], inter["pcs1"][4], inter["f1"][3], inter["f1"][4])
    got = su(cu(inter["pcs1"][3]), cu(inter["pcs1"][4]), cu(inter["f1"][3]), cu(inter["f1"][4]))
    assert float((got.cpu() - want).abs().max()) < 2e-5 * max(1.0, float(want.abs().max()))
    # RefineFlowRegressor / PointConvTransFlowV2 (flow3_r: no flow input)
    u1 = want
    u2 = tp.set_upconv(sd, "su3", 16, inter["pcs2"][3], inter["pcs2"][4], inter["f2"][3], inter["f2"][4])
    rf = M.RefineFlowRegressor(nsample=16, in_channel=256, feat_channel=0, mlp=[256, 256], flow_mlp=[128, 128], use_flow=False)
    rf.load_state_dict(_sub_sd(sd, "flow3_r."), strict=True)
    outs = rf(cu(inter["pcs1"][3]), cu(inter["pcs2"][3]), cu(u1), cu(u2))
    ref = tp.cost_volume(sd, "flow3_r.cost", 16, False, inter["pcs1"][3], inter["pcs2"][3], None, u1, u2)
    for n, w_, g_ in zip(("cost_fwd", "cost_bwd", "feats", "flow"), ref, outs):
        g_ = g_.cpu()
        w_ = w_ if n != "cost_bwd" else torch.nn.functional.pad(w_, (0, g_.shape[2] - w_.shape[2]))
        assert g_.shape == w_.shape and float((g_ - w_).abs().max()) < 3e-5 * max(1.0, float(w_.abs().max())), n
    # UpsampleFlow / PointWarping
    flow3 = inter["l3"][3]
    up = M.UpsampleFlow()(cu(inter["pcs1"][2]), cu(inter["pcs1"][3]), cu(flow3), k=5)
    want_up = tp.upsample_flow(inter["pcs1"][2], inter["pcs1"][3], flow3, k=5)
    assert float((up.cpu() - want_up).abs().max()) < 1e-5
    wp = M.PointWarping()(cu(inter["pcs1"][2]), cu(inter["pcs2"][2]), cu(want_up), 5)
    want_wp = tp.point_warping(inter["pcs1"][2], inter["pcs2"][2], want_up, 5)
    assert float((wp.cpu() - want_wp).abs().max()) < 1e-5
    assert M.PointWarping()(cu(inter["pcs1"][2]), cu(inter["pcs2"][2])) is not None   # flow1=None -> pos2 itself


def test_tflow_four_channel_flow_heads_vs_reference_golden(golden_dir):
    """f-4: the reference's add_Seg_after_FLow = True variant (4-channel flow heads; golden written by the unmodified
    reference with its flag switched on): FPS indices exact, all four flow levels within 1e-4."""
    from ssf_slam_b200.model import TFlow
    from ssf_slam_b200.frontend import SceneFlowFrontEnd
    g = np.load(os.path.join(golden_dir, "tflow_seg4_n2048.npz"))
    net = TFlow(add_seg_after_flow=True)
    net.load_state_dict(tp.random_init_state_dict(int(g["weight_seed"]), 4), strict=True)
    pc1 = torch.from_numpy(g["pos1"].T.copy()).unsqueeze(0).cuda()
    pc2 = torch.from_numpy(g["pos2"].T.copy()).unsqueeze(0).cuda()
    flows, fps = net(pc1, pc2)
    for i in range(3):
        assert np.array_equal(fps[i][0].cpu().numpy(), g["fps%d" % (i + 1)])
    errs = [float(np.abs(flows[i][0].cpu().numpy() - g["flow%d" % i]).max()) for i in range(4)]
    print("4-channel flow max-abs error per level:", errs)
    assert flows[0].shape == (1, 4, 2048) and max(errs) <= FLOW_TOL
    out = SceneFlowFrontEnd(net, tau=0.10).process(g["pos1"][None], g["pos2"][None], return_flow=True)
    assert out["flow"].shape == (1, 2048, 4) and out["mask"].shape == (1, 2048)


def test_tflow_four_channel_input_variant_vs_reference_golden(golden_dir):
    """f-4: the `_afterPC` variant (first layer Conv1d(4, 32), called with [xyz | label] features); golden written by the
    unmodified TFlowV3_Occlussion_addSeg_afterPC.TFlow: FPS indices exact, all four flow levels within 1e-4.  Also the base
    model's feats1 / feats2 arguments: used only when both are given, as in the reference."""
    from ssf_slam_b200.model import TFlow
    from ssf_slam_b200._native import SsfError
    g = np.load(os.path.join(golden_dir, "tflow_afterpc_n2048.npz"))
    net = TFlow(input_channels=4)
    net.load_state_dict(tp.random_init_state_dict(int(g["weight_seed"]), 3, input_channels=4), strict=True)
    pc1 = torch.from_numpy(g["pos1"].T.copy()).unsqueeze(0).cuda()
    pc2 = torch.from_numpy(g["pos2"].T.copy()).unsqueeze(0).cuda()
    f1 = torch.cat([pc1, torch.from_numpy(g["lab1"])[None, None].cuda()], dim=1)
    f2 = torch.cat([pc2, torch.from_numpy(g["lab2"])[None, None].cuda()], dim=1)
    flows, fps = net(pc1, pc2, f1, f2)
    for i in range(3):
        assert np.array_equal(fps[i][0].cpu().numpy(), g["fps%d" % (i + 1)])
    errs = [float(np.abs(flows[i][0].cpu().numpy() - g["flow%d" % i]).max()) for i in range(4)]
    print("4-channel input variant, flow max-abs error per level:", errs)
    assert max(errs) <= FLOW_TOL
    with pytest.raises(SsfError):
        net(pc1, pc2)                      # Conv1d(4, 32) on 3-channel coordinates fails in the reference too
    base = TFlow()
    base.load_state_dict(tp.random_init_state_dict(0), strict=True)
    a, _ = base(pc1, pc2)
    b, _ = base(pc1, pc2, pc1, None)       # one of the two missing -> coordinates are used
    c, _ = base(pc1, pc2, pc1, pc2)        # both given and equal to the coordinates -> same result
    assert torch.equal(a[0], b[0]) and torch.equal(a[0], c[0])


@pytest.mark.parametrize("C,k", [(3, 5), (64, 7), (128, 5), (256, 3), (96, 3), (64, 3)])
def test_upsample_flow_channel_widths(C, k):
    """UpsampleFlow (inverse-distance interpolation) at every channel width of the network: scalar path (C = 3, 96), 8-byte
    (C = 64) and 16-byte (C = 128, 256) channel vectors per lane, against the oracle's restatement of soflow.py:1443-1475."""
    from ssf_slam_b200 import model as M
    g = torch.Generator().manual_seed(C + k)
    xyz = torch.randn(2, 700, 3, generator=g) * 10
    sparse = torch.randn(2, 200, 3, generator=g) * 10
    val = torch.randn(2, 200, C, generator=g)
    got = M.upsample_pm(xyz.cuda(), sparse.cuda(), val.cuda(), k).cpu()
    want = tp.upsample_flow(xyz.transpose(1, 2).contiguous(), sparse.transpose(1, 2).contiguous(), val.transpose(1, 2).contiguous(), k)
    assert got.shape == (2, 700, C)
    assert float((got - want.transpose(1, 2)).abs().max()) < 2e-5


@pytest.mark.parametrize("C,k,mode", [(3, 5, 0), (3, 7, 0), (4, 3, 0), (1, 16, 0), (3, 7, 1), (3, 3, 1), (64, 5, 0), (64, 3, 0),
                                      (96, 3, 0), (128, 7, 0), (32, 3, 0), (8, 8, 0), (64, 16, 0)])
def test_interpolate_kernel_variants_are_bit_identical(C, k, mode):
    """C <= 4 runs one thread per query (layers.cu interpolate_thread_kernel), C <= 128 in 16-byte pieces 8 / 16 / 32 lanes per
    query (interpolate_group_kernel, k <= 8), everything else one warp per query.  Same operations in the same order: the first
    C channels of a zero-padded 256-channel call (warp kernel) must be bit-identical; mode 1 (PointWarping, C == 3 only) is
    checked against query - v of the mode-0 result, clamped."""
    from ssf_slam_b200 import functional as F_
    g = torch.Generator().manual_seed(10 * C + k)
    q = (torch.randn(2, 1111, 3, generator=g) * 10).cuda()
    sp = (torch.randn(2, 300, 3, generator=g) * 10).cuda()
    sp[:, :5] = q[:, :5]      # coincident points: the 1e-10 distance floor
    val = torch.randn(2, 300, C, generator=g).cuda()
    idx = F_.knn_idx(16, q, sp)
    wide = torch.zeros(2, 300, 256, device="cuda")
    wide[..., :C] = val
    ref = F_.interpolate(q, sp, wide, idx, mode=0, clampv=100.0, k=k)[..., :C]
    if mode == 0:
        got = F_.interpolate(q, sp, val, idx, mode=0, clampv=100.0, k=k)
        assert torch.equal(got, ref)
    else:
        got = F_.interpolate(q, sp, val, idx, mode=1, clampv=10.0, k=k)
        unclamped = F_.interpolate(q, sp, wide, idx, mode=0, clampv=3.0e38, k=k)[..., :3]
        assert torch.equal(got, (q - unclamped).clamp(-10.0, 10.0))


@pytest.mark.parametrize("m,P", [(128, 1000), (256, 893), (512, 301), (128, 7)])
def test_attention_mix_persistent_kernel(m, P):
    """The S x S cross attention of the wide cost volumes (cv_parts.cu attention_mix_kernel; soflow.py:420-422,453-458) against
    a plain PyTorch fp32 statement of the same formula; P is chosen so that the persistent CTAs wrap around several times and
    the last round is ragged (both shared-memory buffers and both mbarrier phases are exercised)."""
    from ssf_slam_b200 import functional as F_
    g = torch.Generator().manual_seed(m + P)
    A = (torch.randn(P, 16, m, generator=g) * 0.5).cuda()
    Aw = (torch.randn(P, 16, m, generator=g) * 0.5).cuda()
    Amix, Awmix = F_.attention_mix(A, Aw)
    S = torch.einsum("pic,pjc->pij", A.double(), Aw.double())
    Q = torch.softmax(S, dim=1) * torch.softmax(S, dim=2)
    want_a = A.double() + torch.einsum("pij,pjc->pic", Q, Aw.double())
    want_w = Aw.double() + torch.einsum("pij,pic->pjc", Q, A.double())
    assert float((Amix.double() - want_a).abs().max()) < 2e-5
    assert float((Awmix.double() - want_w).abs().max()) < 2e-5
    a2, w2 = F_.attention_mix(A, Aw)          # deterministic
    assert torch.equal(a2, Amix) and torch.equal(w2, Awmix)


@pytest.mark.parametrize("rows,K,cout", [(5000, 64, 3), (777, 256, 4), (33, 128, 3), (100000, 96, 1), (70001, 64, 3), (66000, 256, 4), (65537, 16, 2),
                                           (4099, 3, 32), (1000, 3, 64), (513, 4, 32)])
def test_linear_narrow_kernels_equal_the_tiled_kernel(rows, K, cout):
    """Flow heads (cout <= 4) and the first per-point layer (K <= 4) run dedicated kernels (layers.cu linear_narrow_*) with the
    same fmaf chain as the 64 x 64 tiled kernel.  A second, all-zero input block forces the tiled kernel (fmaf(0, w, acc) == acc),
    so both results must be bit-identical -- bias, activation, both clamps and the residual included."""
    from ssf_slam_b200 import functional as F_
    g = torch.Generator().manual_seed(rows + K + cout)
    x = torch.randn(rows, K, generator=g).cuda()
    Wt = (torch.randn(K + 16, cout, generator=g) / K ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    add = (torch.randn(rows, cout, generator=g) * 30).cuda()
    z = torch.zeros(rows, 16, device="cuda")
    for kw in (dict(bias=b, act=2), dict(bias=b, clamp1=1.5, add=add, clamp2=40.0), dict()):
        got = F_.linear(x, Wt, cout, 0, **kw)
        ref = F_.linear(x, Wt, cout, 0, z, K, **kw)
        assert torch.equal(got, ref), kw.keys()
    want = x.double() @ Wt[:K].double() + b.double()
    got = F_.linear(x, Wt, cout, 0, bias=b)
    assert float((got.double() - want).abs().max()) < 1e-4
