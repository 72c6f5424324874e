"""B200-native ActiveSceneFlow ``TFlow`` with the reference's call surface.

``TFlow(npoint=8192)`` / ``forward(pc1[B,3,N], pc2[B,3,N]) -> ([flow B x3xN, B x3x2048, B x3x512, B x3x256],
[fps idx B x2048, B x512, B x256])`` exactly as ``ASF/TFlowV3_Occlussion.py:65-196``, and the module tree exposes
the same ``state_dict`` keys (317 tensors), so the reference's checkpoint (``model.best.t7``, loaded at
``ASF/main_sju_occ_ros.py:698-711``, also under ``nn.DataParallel``'s ``module.`` prefix) loads with
``strict=True``.  The torch modules below are parameter holders only: ``forward`` never runs a torch op on the
hot path; it folds eval-mode BatchNorm, re-lays the weights K-major once, and drives the fused sm_100a kernels
(``ssf_slam_b200.functional``).  Inference / eval mode only, as in the reference's ROS drivers.

Structure of the forward (what the reference does per level, re-expressed point-major and fused):
  pyramid   : FPS -> kNN -> [gather -> 3x(conv+BN+ReLU) -> max]           ASF/utils/utils.py:208-248
  up-conv   : kNN -> [gather -> 2x(conv+BN+ReLU) -> max] -> 2x linear      ASF/utils/utils.py:274-315
  regressor : warp -> 2x kNN -> fused cost volume -> segmented softmax/sum -> [gather -> mlp_convs4 -> max]
              -> flow head                                                  ASF/utils/soflow.py:354-525,1223-1257
  upsample  : kNN / 3-NN -> inverse-distance interpolation                  ASF/utils/soflow.py:1443-1475
Both clouds of a pair travel through the pyramid and the up-convs as one batch of 2B clouds.
"""
import torch
import torch.nn as nn

from . import functional as F_
from . import _native as nat
from . import tc

ACT_NONE, ACT_RELU, ACT_LEAKY = F_.ACT_NONE, F_.ACT_RELU, F_.ACT_LEAKY


# --------------------------------------------------------------------------- parameter holders
# (names mirror the reference module tree so that state_dict keys are identical)

class _OwnWeights:
    """Layer objects used on their own (the B-layer boundary: reference TFlow code constructing our classes) prepare
    the kernel operands from their own parameters, once per device / load_state_dict."""
    _kind = None

    def _own(self, device):
        cache = getattr(self, "_own_cache", None)
        if cache is None or cache[0] != device:
            cache = (device, prepare_block(self._kind, self.state_dict(), device))
            object.__setattr__(self, "_own_cache", cache)
        return cache[1]

    def _load_from_state_dict(self, *args, **kwargs):
        object.__setattr__(self, "_own_cache", None)
        return super()._load_from_state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True, **kw):
        return super().load_state_dict(_strip_module_prefix(state_dict), strict=strict, **kw)


def _strip_module_prefix(state_dict):
    """nn.DataParallel checkpoints: drop a leading ``module.`` when every key carries it."""
    if len(state_dict) and all(k.startswith("module.") for k in state_dict):
        return type(state_dict)((k[7:], v) for k, v in state_dict.items())
    return state_dict


def _pm(x):
    """reference layout [B,C,N] -> point-major [B,N,C] (contiguous fp32)"""
    return F_.transpose(x.contiguous().float())


class _LeakyConv1d(nn.Module):
    """`Conv1d` blocks: `.composed_module.0` is the conv (ASF/TFlowV3_Occlussion.py:22-38 no bias;
    ASF/utils/soflow.py:1260-1276 with bias)."""

    def __init__(self, cin, cout, bias):
        super().__init__()
        self.composed_module = nn.Sequential(nn.Conv1d(cin, cout, 1, bias=bias), nn.Identity(), nn.Identity())


class PointNetSetAbstraction(_OwnWeights, nn.Module):
    """ASF/utils/utils.py:185-248 (radius / group_all are accepted and ignored, as there: the layer groups by kNN)."""
    _kind = "sa"

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all=False):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all = npoint, radius, nsample, group_all
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel + 3
        for c in mlp:
            self.mlp_convs.append(nn.Conv2d(last, c, 1, bias=False))
            self.mlp_bns.append(nn.BatchNorm2d(c))
            last = c
        self.eval()

    @torch.no_grad()
    def forward(self, xyz, points):
        """xyz [B,3,N], points [B,D,N] -> (new_xyz [B,3,S], new_points [B,mlp[-1],S], fps_idx i32 [B,S])
        (ASF/utils/utils.py:208-248; eval mode)."""
        nat.require_device()
        F_.knn_cache_clear()
        nx, nf, fi = set_abstraction_pm(self._own(xyz.device), self.npoint, self.nsample, _pm(xyz), _pm(points))
        F_.knn_cache_clear()
        return F_.transpose(nx), F_.transpose(nf), fi


class PointNetSetUpConv(_OwnWeights, nn.Module):
    """ASF/utils/utils.py:250-315."""
    _kind = "su"

    def __init__(self, nsample, radius, f1_channel, f2_channel, mlp, mlp2, knn=True):
        super().__init__()
        self.nsample, self.radius, self.knn = nsample, radius, knn
        self.mlp1_convs = nn.ModuleList()
        self.mlp2_convs = nn.ModuleList()
        last = f2_channel + 3
        for c in mlp:
            self.mlp1_convs.append(nn.Sequential(nn.Conv2d(last, c, 1, bias=False), nn.BatchNorm2d(c), nn.ReLU()))
            last = c
        last = (mlp[-1] if len(mlp) else last) + f1_channel
        for c in mlp2:
            self.mlp2_convs.append(nn.Sequential(nn.Conv1d(last, c, 1, bias=False), nn.BatchNorm1d(c), nn.ReLU()))
            last = c
        self.eval()

    @torch.no_grad()
    def forward(self, pos1, pos2, feature1, feature2):
        """pos1 [B,3,N1] dense, pos2 [B,3,N2] sparse, feature1 [B,C1,N1], feature2 [B,C2,N2] -> [B,mlp2[-1],N1]
        (ASF/utils/utils.py:274-315; eval mode)."""
        nat.require_device()
        F_.knn_cache_clear()
        out = set_upconv_pm(self._own(pos1.device), self.nsample, _pm(pos1), _pm(pos2), _pm(feature1), _pm(feature2))
        F_.knn_cache_clear()
        return F_.transpose(out)


class PointWarping(nn.Module):
    """ASF/utils/soflow.py:1222-1257: pos2 pulled back by the flow interpolated from pos1 + flow1 (no parameters)."""

    @torch.no_grad()
    def forward(self, pos1, pos2, flow1=None, nsample=None):
        if flow1 is None:
            return pos2
        nat.require_device()
        F_.knn_cache_clear()
        out = point_warping_pm(_pm(pos1), _pm(pos2), _pm(flow1), nsample)
        F_.knn_cache_clear()
        return F_.transpose(out)


class UpsampleFlow(nn.Module):
    """ASF/utils/soflow.py:1442-1475: normalised inverse-distance interpolation from the sparse level (no parameters)."""

    @torch.no_grad()
    def forward(self, xyz, sparse_xyz, sparse_flow, k=3):
        nat.require_device()
        F_.knn_cache_clear()
        out = upsample_pm(_pm(xyz), _pm(sparse_xyz), _pm(sparse_flow), k)
        F_.knn_cache_clear()
        return F_.transpose(out)


class PointConvTransFlowV2(_OwnWeights, nn.Module):
    """ASF/utils/soflow.py:281-525 (bn=False, 3-channel flow head)."""
    _kind = "cv"

    def __init__(self, nsample, in_channel, sf_channel, mlp, flow_mlp, use_flow=True, flow_channels=3):
        super().__init__()
        self.nsample, self.use_flow = nsample, use_flow
        self.in_channel, self.sf_channel, self.mlp, self.flow_mlp = in_channel, sf_channel, list(mlp), list(flow_mlp)
        m = mlp[-1]

        def stack(cin):
            ml = nn.ModuleList()
            for c in mlp:
                ml.append(nn.Conv2d(cin, c, 1))
                cin = c
            return ml

        self.mlp_convs = stack(in_channel * 2)
        self.mlp_convs2 = stack(in_channel * 2)
        self.weightnet1 = nn.Sequential(nn.Conv2d(m, m, 1, bias=False), nn.BatchNorm2d(m), nn.ReLU(),
                                        nn.Conv2d(m, m // 2, 1, bias=False), nn.BatchNorm2d(m // 2), nn.ReLU(),
                                        nn.Conv2d(m // 2, 1, 1))
        self.mlp_convs3 = stack(m + sf_channel + 3)
        self.mlp_convs4 = stack(m * 2 + sf_channel + 3)
        self.flow_mlp_convs = nn.ModuleList()
        last = m
        for c in flow_mlp:
            self.flow_mlp_convs.append(_LeakyConv1d(last, c, bias=True))
            last = c
        self.fc = nn.Conv1d(last, flow_channels, 1)   # 4 with the reference's add_Seg_after_FLow flag (soflow.py:343-346)
        self.eval()

    @torch.no_grad()
    def forward(self, xyz1, xyz2, xyz2w, points1, points2, sf=None, sf_feat=None):
        """xyz1 [B,3,N1], xyz2 [B,3,N2], xyz2w [B,3,N2] or None, points1 [B,D,N1], points2 [B,D,N2], sf [B,3,N1],
        sf_feat [B,F,N1] -> (cost_fwd [B,m,N1], cost_bwd [B,m,N2], feats [B,flow_mlp[-1],N1], flow [B,3,N1])
        (ASF/utils/soflow.py:354-525).  The backward cost has all N2 columns (zeros where the reference's tensor ends)."""
        nat.require_device()
        if self.nsample != 16:
            raise nat.SsfError("PointConvTransFlowV2: the fused kernels are built for nsample == 16 (as in TFlow)")
        F_.knn_cache_clear()
        outs = cost_volume_pm(self._own(xyz1.device), _pm(xyz1), _pm(xyz2), None if xyz2w is None else _pm(xyz2w), _pm(points1),
                              None, _pm(points2), None, sf=None if sf is None else _pm(sf),
                              sf_feat=None if sf_feat is None else _pm(sf_feat))
        F_.knn_cache_clear()
        return tuple(F_.transpose(o) for o in outs)


class RefineFlowRegressor(nn.Module):
    """ASF/TFlowV3_Occlussion.py:41-62: optional PointWarping, then the cost volume."""

    def __init__(self, nsample=8, in_channel=128, feat_channel=128, mlp=(128, 128, 128), flow_mlp=(128, 128), use_flow=True,
                 flow_channels=3):
        super().__init__()
        self.use_flow, self.nsample = use_flow, nsample
        self.cost = PointConvTransFlowV2(nsample, in_channel, feat_channel, list(mlp), list(flow_mlp), use_flow=use_flow,
                                         flow_channels=flow_channels)
        self.warping = PointWarping()

    @torch.no_grad()
    def forward(self, pc1, pc2, feats1, feats2, wraping_num=5, c_flow=None, flow_feats=None):
        pc2_warp = None if c_flow is None else self.warping(pc1, pc2, c_flow, wraping_num)
        if self.use_flow:
            return self.cost(pc1, pc2, pc2_warp, feats1, feats2, c_flow, flow_feats)
        return self.cost(pc1, pc2, pc2_warp, feats1, feats2)


# ------------------------------------------------------------------------------ weight preparation

def _fold_bn(w2d, bn, prefix, sd):
    """conv (no bias) + eval BatchNorm -> (W', b') with W' [Cout, Cin]."""
    g, b = sd[prefix + ".weight"].double(), sd[prefix + ".bias"].double()
    mu, var = sd[prefix + ".running_mean"].double(), sd[prefix + ".running_var"].double()
    scale = g / torch.sqrt(var + bn)
    return (w2d.double() * scale[:, None]).float(), (b - mu * scale).float()


def _kmajor(w2d):
    return w2d.t().contiguous()


def prepare_weights(state_dict, device):
    """Reference-format state_dict -> dict of K-major, BN-folded fp32 device tensors used by the kernels."""
    return _prepare(state_dict, device, None)


def prepare_block(kind, state_dict, device):
    """Operands of ONE layer object from its own state_dict: kind in {"sa", "su", "cv"} (PointNetSetAbstraction,
    PointNetSetUpConv, PointConvTransFlowV2).  Used by the drop-in layer classes' own ``forward``."""
    sd = {"x." + k: v for k, v in state_dict.items()}
    return _prepare(sd, device, kind)


def _prepare(state_dict, device, only):
    sd = {k[7:] if k.startswith("module.") else k: v.detach().to("cpu") for k, v in state_dict.items()}
    W = {}
    eps = 1e-5

    def w2(key):
        w = sd[key]
        return w.reshape(w.shape[0], w.shape[1]).float()

    def dev(t):
        return t.contiguous().to(device)

    def img(w_nk):
        """[Cout, Cin] -> tensor-core weight image (ssf_dense_tc)"""
        return dev(tc.dense_image(w_nk.contiguous()))

    if only is None:
        W["pc0"] = dev(_kmajor(w2("point_conv.0.composed_module.0.weight")))
        W["pc1"] = dev(_kmajor(w2("point_conv.1.composed_module.0.weight")))
        W["pc1_img"] = img(w2("point_conv.1.composed_module.0.weight"))
        for name in ("deconv3_2", "deconv2_1", "deconv1_0"):
            W[name] = dev(_kmajor(w2(name + ".composed_module.0.weight")))
            W[name + "_img"] = img(w2(name + ".composed_module.0.weight"))

    for name in (("sa1", "sa2", "sa3", "sa4") if only is None else (("x",) if only == "sa" else ())):
        layers = []
        i = 0
        while "%s.mlp_convs.%d.weight" % (name, i) in sd:
            w, b = _fold_bn(w2("%s.mlp_convs.%d.weight" % (name, i)), eps, "%s.mlp_bns.%d" % (name, i), sd)
            layers.append((w, b))
            i += 1
        assert len(layers) == 3
        w1, b1 = layers[0]
        W[name] = dict(Wd=dev(_kmajor(w1[:, :3])), Wg=dev(_kmajor(w1[:, 3:])), b1=dev(b1),
                       W2=dev(_kmajor(layers[1][0])), b2=dev(layers[1][1]), C2=layers[1][0].shape[0],
                       W3=dev(_kmajor(layers[2][0])), b3=dev(layers[2][1]), C3=layers[2][0].shape[0], C1=w1.shape[0],
                       Wg_img=img(w1[:, 3:]), W2_img=img(layers[1][0]), W3_img=img(layers[2][0]))

    for name in (("su3", "su2", "su1", "su0") if only is None else (("x",) if only == "su" else ())):
        m1 = []
        i = 0
        while "%s.mlp1_convs.%d.0.weight" % (name, i) in sd:
            m1.append(_fold_bn(w2("%s.mlp1_convs.%d.0.weight" % (name, i)), eps, "%s.mlp1_convs.%d.1" % (name, i), sd))
            i += 1
        m2 = []
        i = 0
        while "%s.mlp2_convs.%d.0.weight" % (name, i) in sd:
            m2.append(_fold_bn(w2("%s.mlp2_convs.%d.0.weight" % (name, i)), eps, "%s.mlp2_convs.%d.1" % (name, i), sd))
            i += 1
        assert len(m1) == 2 and len(m2) == 2
        w1, b1 = m1[0]
        c2 = w1.shape[1] - 3
        W[name] = dict(Wg=dev(_kmajor(w1[:, :c2])), Wd=dev(_kmajor(w1[:, c2:])), b1=dev(b1), C1=w1.shape[0],
                       W2=dev(_kmajor(m1[1][0])), b2=dev(m1[1][1]), C2=m1[1][0].shape[0],
                       M1=dev(_kmajor(m2[0][0])), mb1=dev(m2[0][1]), MC1=m2[0][0].shape[0],
                       M2=dev(_kmajor(m2[1][0])), mb2=dev(m2[1][1]), MC2=m2[1][0].shape[0],
                       Wg_img=img(w1[:, :c2]), W2_img=img(m1[1][0]), M1_img=img(m2[0][0]), M2_img=img(m2[1][0]))

    for name in (("flow3_r", "flow2_r", "flow1_r", "flow0_r") if only is None else (("x",) if only == "cv" else ())):
        p = name + ".cost" if only is None else name
        wa, ba = w2(p + ".mlp_convs.0.weight"), sd[p + ".mlp_convs.0.bias"].float()
        ww, bw = w2(p + ".mlp_convs2.0.weight"), sd[p + ".mlp_convs2.0.bias"].float()
        m, D = wa.shape[0], wa.shape[1] // 2
        w3, b3 = w2(p + ".mlp_convs3.0.weight"), sd[p + ".mlp_convs3.0.bias"].float()
        Fc = w3.shape[1] - m - 3
        w4, b4 = w2(p + ".mlp_convs4.0.weight"), sd[p + ".mlp_convs4.0.bias"].float()
        wn1, bn1 = _fold_bn(w2(p + ".weightnet1.0.weight"), eps, p + ".weightnet1.1", sd)
        wn2, bn2 = _fold_bn(w2(p + ".weightnet1.3.weight"), eps, p + ".weightnet1.4", sd)
        d = dict(m=m, D=D, F=Fc,
                 Wab=dev(torch.cat([_kmajor(wa), _kmajor(ww)], dim=1)),       # [2D, 2m]; rows [0,D) per-query, [D,2D) gathered
                 bab=dev(torch.cat([ba, bw])),
                 W2a=dev(_kmajor(w2(p + ".mlp_convs.1.weight"))), b2a=dev(sd[p + ".mlp_convs.1.bias"].float()),
                 W2w=dev(_kmajor(w2(p + ".mlp_convs2.1.weight"))), b2w=dev(sd[p + ".mlp_convs2.1.bias"].float()),
                 W3=dev(_kmajor(w3)), b3=dev(b3),                              # rows [0,m) A | [m,m+F) sf_feat | [m+F,m+F+3) dir
                 W3b=dev(_kmajor(w2(p + ".mlp_convs3.1.weight"))), b3b=dev(sd[p + ".mlp_convs3.1.bias"].float()),
                 Wn1=dev(_kmajor(wn1)), bn1=dev(bn1), Wn2=dev(_kmajor(wn2)), bn2=dev(bn2),
                 wn3=dev(w2(p + ".weightnet1.6.weight").reshape(-1)), bn3=float(sd[p + ".weightnet1.6.bias"].reshape(-1)[0]),
                 W4=dev(_kmajor(w4)), b4=dev(b4),                              # rows [0,m) fwd | [m,2m) bwd | [2m,2m+F) sf_feat | dir
                 W42=dev(_kmajor(w2(p + ".mlp_convs4.1.weight"))), b42=dev(sd[p + ".mlp_convs4.1.bias"].float()),
                 fc=dev(_kmajor(w2(p + ".fc.weight"))), fcb=dev(sd[p + ".fc.bias"].float()))
        if m == 64:  # tensor-core path (csrc/cost_volume_tc.cu): split TF32 weight images
            blob, par = tc.cost_volume_tc_pack(
                w2(p + ".mlp_convs.1.weight"), w2(p + ".mlp_convs2.1.weight"), w3[:, :m], w2(p + ".mlp_convs3.1.weight"),
                wn1, wn2, sd[p + ".mlp_convs.1.bias"].float(), sd[p + ".mlp_convs2.1.bias"].float(),
                sd[p + ".mlp_convs3.1.bias"].float(), bn1, bn2, w2(p + ".weightnet1.6.weight").reshape(-1),
                _kmajor(w3)[m + Fc:], float(sd[p + ".weightnet1.6.bias"].reshape(-1)[0]))
            d["tc_blob"], d["tc_par"] = dev(blob), dev(par)
        w42 = w2(p + ".mlp_convs4.1.weight")
        d.update(Hab_img=img(torch.cat([wa[:, :D], ww[:, :D]], dim=0)), Gab_img=img(torch.cat([wa[:, D:], ww[:, D:]], dim=0)),
                 W2a_img=img(w2(p + ".mlp_convs.1.weight")), W2w_img=img(w2(p + ".mlp_convs2.1.weight")),
                 W3a_img=img(w3[:, :m]), W3b_img=img(w2(p + ".mlp_convs3.1.weight")), Wn1_img=img(wn1), Wn2_img=img(wn2),
                 Hp_img=img(torch.cat([w4[:, :m], w4[:, 2 * m:2 * m + Fc]], dim=1)), G4_img=img(w4[:, m:2 * m]), W42_img=img(w42))
        if Fc > 0:
            d["H3_img"] = img(w3[:, m:m + Fc])
        d["W3a"] = d["W3"][:m]
        d["W3d"] = d["W3"][m + Fc:]
        d["W4d"] = d["W4"][2 * m + Fc:]
        fm = []
        i = 0
        while "%s.flow_mlp_convs.%d.composed_module.0.weight" % (p, i) in sd:
            k = "%s.flow_mlp_convs.%d.composed_module.0" % (p, i)
            fm.append((dev(_kmajor(w2(k + ".weight"))), dev(sd[k + ".bias"].float()), sd[k + ".weight"].shape[0], img(w2(k + ".weight"))))
            i += 1
        d["flow_mlp"] = fm
        W[name] = d
    return W if only is None else W["x"]


# ------------------------------------------------------------------------------ point-major forward pieces

def _tc():
    return F_.USE_TC


def set_abstraction_pm(w, npoint, nsample, xyz, feats):
    """xyz [B,N,3], feats [B,N,D] -> (new_xyz [B,S,3], new_feats [B,S,C3], fps_idx [B,S])."""
    fps_idx = F_.fps(xyz, npoint)
    new_xyz = F_.gather_rows(xyz, fps_idx)
    idx = F_.knn_idx(nsample, new_xyz, xyz)
    D = feats.shape[-1]
    if _tc():
        # first conv split per point (Wg.feats once per source point), rows (n,s) formed on the fly inside the second
        # conv's A operand, third conv + max over nsample in the epilogue: the grouped tensor never exists in HBM
        G = F_.dense_tc(w["Wg_img"], w["C1"], D, x1=feats)
        x2 = F_.dense_tc(w["W2_img"], w["C2"], w["C1"], G=G, b1=w["b1"], Wd1=w["Wd"], act1=ACT_RELU, idx=idx, pos_src=xyz,
                         pos_q=new_xyz, bias=w["b2"], act=ACT_RELU)
        out = F_.dense_tc(w["W3_img"], w["C3"], w["C2"], x1=x2, S=nsample, bias=w["b3"], act=ACT_RELU, epi=F_.EPI_MAX)
        return new_xyz, out.view(xyz.shape[0], npoint, w["C3"]), fps_idx
    G = F_.linear(feats, w["Wg"], w["C1"])
    out = F_.group_mlp_max(G, idx, xyz, new_xyz, w["Wd"], w["b1"], w["W2"], w["b2"], w["C2"], w["W3"], w["b3"], w["C3"],
                           act=ACT_RELU)
    return new_xyz, out, fps_idx


def set_upconv_pm(w, nsample, pos1, pos2, feat1, feat2, return_idx=False):
    """Feature propagation sparse (pos2, feat2) -> dense (pos1, feat1): [B,N1,mlp2[-1]]."""
    idx = F_.knn_idx(nsample, pos1, pos2)
    out = _set_upconv_body(w, idx, pos1, pos2, feat1, feat2)
    return (out, idx) if return_idx else out


def _set_upconv_body(w, idx, pos1, pos2, feat1, feat2):
    if _tc():
        G = F_.dense_tc(w["Wg_img"], w["C1"], feat2.shape[-1], x1=feat2)
        pooled = F_.dense_tc(w["W2_img"], w["C2"], w["C1"], G=G, b1=w["b1"], Wd1=w["Wd"], act1=ACT_RELU, idx=idx, pos_src=pos2,
                             pos_q=pos1, bias=w["b2"], act=ACT_RELU, epi=F_.EPI_MAX)
        x = F_.dense_tc(w["M1_img"], w["MC1"], w["C2"] + feat1.shape[-1], x1=pooled, x2=feat1, bias=w["mb1"], act=ACT_RELU)
        return F_.dense_tc(w["M2_img"], w["MC2"], w["MC1"], x1=x, bias=w["mb2"], act=ACT_RELU)
    G = F_.linear(feat2, w["Wg"], w["C1"])
    pooled = F_.group_mlp_max(G, idx, pos2, pos1, w["Wd"], w["b1"], w["W2"], w["b2"], w["C2"], act=ACT_RELU)
    x = F_.linear(pooled, w["M1"], w["MC1"], 0, feat1, w["C2"], bias=w["mb1"], act=ACT_RELU)
    return F_.linear(x, w["M2"], w["MC2"], bias=w["mb2"], act=ACT_RELU)


def upsample_pm(xyz, sparse_xyz, sparse_val, k=3, idx=None):
    """UpsampleFlow: [B,S,C] on sparse_xyz -> [B,N,C] on xyz.  `idx` may be any (distance, index)-ordered neighbour list
    of xyz in sparse_xyz with at least k entries per row (its k-prefix IS the k-NN list the reference computes)."""
    if idx is None:
        idx = F_.knn_idx(k, xyz, sparse_xyz)
    return F_.interpolate(xyz, sparse_xyz, sparse_val, idx, mode=0, clampv=100.0, k=k)


def point_warping_pm(pos1, pos2, flow1, k):
    """PointWarping: pos2 pulled back by the flow interpolated from pos1 + flow1 (positions clamped to +-10 m)."""
    if flow1.shape[-1] != 3:   # add_Seg_after_FLow: the 4th channel (segmentation logit) does not move points (soflow.py:1228,1252)
        flow1 = flow1[..., :3].contiguous()
    moved = pos1 + flow1  # one rounded fp32 add per coordinate, as `pos1 + flow1` at soflow.py:1231 (glue, 3 floats/point)
    idx = F_.knn_idx(3 if k is None else k, pos2, moved)
    return F_.interpolate(pos2, moved, flow1, idx, mode=1, clampv=10.0)


def _cost_volume_wide_tc(w, Gab, Hab, H3, xyz1, xyz2, idx, idxw, m):
    """m >= 128 (levels 2, 3): the dense layers as tensor-core GEMMs (rows = (point, neighbour)), the S x S attention and
    the softmax-weighted forward cost on CUDA cores between them.  Same outputs as F_.cost_volume."""
    B, N1, _ = xyz1.shape
    L = ACT_LEAKY
    br = []
    for b_, (ii, w2i, b2) in enumerate(((idx, w["W2a_img"], w["b2a"]), (idxw, w["W2w_img"], w["b2w"]))):
        br.append(F_.dense_tc(w2i, m, m, G=Gab, offG=b_ * m, H=Hab, offH=b_ * m, act1=L, idx=ii, bias=b2, act=L))
    A, Aw = br                                                                  # [B,N1,16,m]
    Amix, Awmix = F_.attention_mix(A, Aw)
    outs = []
    for rows, ii in ((A, idx), (Aw, idxw)):
        c1 = F_.dense_tc(w["W3a_img"], m, m, x1=rows, idx=ii, pos_src=xyz2, pos_q=xyz1, Hq=H3, Wd2=w["W3d"], act=L)
        outs.append(F_.dense_tc(w["W3b_img"], m, m, x1=c1, bias=w["b3b"], act=L))
    C, Cw = outs
    logits = []
    for rows in (Amix, Awmix):
        t1 = F_.dense_tc(w["Wn1_img"], m, m, x1=rows, bias=w["bn1"], act=ACT_RELU)
        logits.append(F_.dense_tc(w["Wn2_img"], m // 2, m, x1=t1, bias=w["bn2"], act=ACT_RELU, epi=F_.EPI_DOT, wvec=w["wn3"],
                                  b0=w["bn3"]))
    g, gw = logits                                                              # [B,N1,16]
    cost_fwd, cost_fwd_cm = F_.softmax_pool(g, C)
    return cost_fwd, cost_fwd_cm, gw.view(B, N1 * 16), Cw.view(B, N1 * 16, m)


def cost_volume_pm(w, xyz1, xyz2, xyz2w, f1a, f1b, f2a, f2b, sf=None, sf_feat=None):
    """PointConvTransFlowV2 on point-major inputs.  f1 = cat[f1a | f1b], f2 = cat[f2a | f2b] (b parts may be None).
    Returns (cost_fwd [B,N1,m], cost_bwd [B,N2,m], feats [B,N1,flow_mlp[-1]], flow [B,N1,3])."""
    m, D, Fc = w["m"], w["D"], w["F"]
    B, N1, _ = xyz1.shape
    N2 = xyz2.shape[1]
    ca = f1a.shape[-1]
    tcp = _tc()
    sf3 = sf if sf is None or sf.shape[-1] == 3 else sf[..., :3].contiguous()   # soflow.py:386-389
    idx = F_.knn_idx(16, xyz1, xyz2, offset=sf3)
    idxw = F_.knn_idx(16, xyz1, xyz2 if xyz2w is None else xyz2w)
    if tcp:
        Hab = F_.dense_tc(w["Hab_img"], 2 * m, D, x1=f1a, x2=f1b, bias=w["bab"])
        Gab = F_.dense_tc(w["Gab_img"], 2 * m, D, x1=f2a, x2=f2b)
    else:
        Hab = F_.linear(f1a, w["Wab"], 2 * m, 0, f1b, ca, bias=w["bab"])
        Gab = F_.linear(f2a, w["Wab"], 2 * m, D, f2b, D + ca)
    if Fc > 0:
        H3 = F_.dense_tc(w["H3_img"], m, Fc, x1=sf_feat, bias=w["b3"]) if tcp else F_.linear(sf_feat, w["W3"], m, m, bias=w["b3"])
    else:
        H3 = w["b3"].view(1, 1, m).expand(B, N1, m).contiguous()
    if tcp and m != 64:
        cost_fwd, cost_fwd_cm, gw, Cw = _cost_volume_wide_tc(w, Gab, Hab, H3, xyz1, xyz2, idx, idxw, m)
    else:
        cost_fwd, cost_fwd_cm, gw, Cw = F_.cost_volume(Gab, Hab, w, H3, xyz1, xyz2, idx, idxw, m)
    csr = F_.build_csr(idxw.view(B, N1 * 16), N2)
    cost_bwd = F_.segment_softmax_sum(gw, Cw, csr, N2)
    # mlp_convs4 input = cat[scrambled fwd | bwd[idx] | sf_feat | dir]; the "scramble" is the reference's .view of
    # the channel-major forward cost as [N1, m] rows (soflow.py:490): reinterpret, do not transpose
    scr = cost_fwd_cm.view(B, N1, m)
    if tcp:
        Hp = F_.dense_tc(w["Hp_img"], m, m + Fc, x1=scr, x2=sf_feat if Fc > 0 else None, bias=w["b4"])
        G4 = F_.dense_tc(w["G4_img"], m, m, x1=cost_bwd)
        x = F_.dense_tc(w["W42_img"], m, m, G=G4, H=Hp, Wd1=w["W4d"], act1=ACT_LEAKY, idx=idx, pos_src=xyz2, pos_q=xyz1,
                        bias=w["b42"], act=ACT_LEAKY, epi=F_.EPI_MAX)
        for Wt, b, c, im in w["flow_mlp"]:
            x = F_.dense_tc(im, c, x.shape[-1], x1=x, bias=b, act=ACT_LEAKY)
    else:
        if Fc > 0:
            Hp = F_.linear(scr, w["W4"], m, 0, sf_feat, 2 * m, bias=w["b4"])
        else:
            Hp = F_.linear(scr, w["W4"], m, 0, bias=w["b4"])
        G4 = F_.linear(cost_bwd, w["W4"], m, m)
        x = F_.group_mlp_max(G4, idx, xyz2, xyz1, w["W4d"], None, w["W42"], w["b42"], m, H=Hp, act=ACT_LEAKY)
        for Wt, b, c, im in w["flow_mlp"]:
            x = F_.linear(x, Wt, c, bias=b, act=ACT_LEAKY)
    flow = F_.linear(x, w["fc"], w["fc"].shape[1], bias=w["fcb"], clamp1=50.0, add=sf, clamp2=50.0)
    return cost_fwd, cost_bwd, x, flow


# ------------------------------------------------------------------------------------- the model

class TFlow(nn.Module):
    """Drop-in for ``TFlowV3_Occlussion.TFlow`` (same ctor, forward signature, outputs and state_dict keys)."""

    SA = (("sa1", 2048, 16), ("sa2", 512, 16), ("sa3", 256, 16), ("sa4", 128, 8))

    def __init__(self, npoint=8192, add_seg_after_flow=False, input_channels=3):
        """``add_seg_after_flow`` is the reference's source-level flag ``add_Seg_after_FLow`` (utils/datasets/carla.py:9):
        4-channel flow heads whose last channel is a segmentation logit; shipped default False.
        ``input_channels=4`` is ``TFlowV3_Occlussion_addSeg_afterPC.TFlow`` (its only difference: ``Conv1d(4, 32)`` as the first
        layer, ASF/TFlowV3_Occlussion_addSeg_afterPC.py:68), fed through ``forward(pc1, pc2, feats1, feats2)`` with
        ``[B,4,N]`` features (xyz + a per-point label)."""
        super().__init__()
        fc_ch = 4 if add_seg_after_flow else 3
        self.input_channels = input_channels
        self.point_conv = nn.Sequential(_LeakyConv1d(input_channels, 32, bias=False), _LeakyConv1d(32, 32, bias=False))
        self.sa1 = PointNetSetAbstraction(2048, 0.5, 16, 32, [32, 32, 64])
        self.sa2 = PointNetSetAbstraction(512, 2.0, 16, 64, [64, 64, 128])
        self.sa3 = PointNetSetAbstraction(256, 4.0, 16, 128, [128, 128, 256])
        self.sa4 = PointNetSetAbstraction(128, 8.0, 8, 256, [256, 256, 512])
        self.su3 = PointNetSetUpConv(16, 2.4, 256, 512, [256, 256], [256, 256])
        self.flow3_r = RefineFlowRegressor(16, 256, 0, [256, 256], [128, 128], use_flow=False, flow_channels=fc_ch)
        self.su2 = PointNetSetUpConv(16, 2.4, 128, 256, [128, 128], [128, 128])
        self.flow2_r = RefineFlowRegressor(16, 128 + 64, 128, [128, 128], [128, 128], flow_channels=fc_ch)
        self.su1 = PointNetSetUpConv(16, 2.4, 64, 128, [64, 64], [64, 64])
        self.flow1_r = RefineFlowRegressor(16, 64 + 32, 128, [64, 64], [64, 64], flow_channels=fc_ch)
        self.su0 = PointNetSetUpConv(16, 2.4, 32, 64, [64, 64], [64, 64])
        self.flow0_r = RefineFlowRegressor(16, 64 + 32, 64, [64, 64], [64, 64], flow_channels=fc_ch)
        self.deconv3_2 = _LeakyConv1d(256, 64, bias=False)
        self.deconv2_1 = _LeakyConv1d(128, 32, bias=False)
        self.deconv1_0 = _LeakyConv1d(64, 32, bias=False)
        self._prepared = None
        self.eval()

    def _load_from_state_dict(self, *args, **kwargs):
        self._prepared = None
        return super()._load_from_state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts the reference's checkpoints as they are saved: plain keys, or every key under nn.DataParallel's ``module.``
        prefix (the reference driver builds such dicts, ASF/main_sju_occ_ros.py:706-709)."""
        self._prepared = None
        return super().load_state_dict(_strip_module_prefix(state_dict), strict=strict, **kw)

    def weights(self, device):
        if self._prepared is None or self._prepared[0] != device:
            self._prepared = (device, prepare_weights(self.state_dict(), device))
        return self._prepared[1]

    @torch.no_grad()
    def forward_pm(self, xyz1, xyz2, f1=None, f2=None):
        """xyz1, xyz2 f32 [B,N,3] (point-major, CUDA) -> (flows pm [[B,N,3],[B,2048,3],[B,512,3],[B,256,3]], fps idx x3).
        f1, f2 f32 [B,N,input_channels]: the optional input features of the reference's forward (TFlowV3_Occlussion.py:111-116:
        ``point_conv`` runs on them instead of the coordinates when BOTH are given)."""
        nat.require_device()
        F_.knn_cache_clear()
        W = self.weights(xyz1.device)
        B = xyz1.shape[0]
        xyz = [torch.cat([xyz1, xyz2], dim=0).contiguous()]  # both clouds as one batch of 2B
        x0 = xyz[0] if f1 is None or f2 is None else torch.cat([f1, f2], dim=0).contiguous()
        if x0.shape[-1] != self.input_channels:
            raise nat.SsfError("point_conv expects %d input channels, got %d" % (self.input_channels, x0.shape[-1]))
        x = F_.linear(x0, W["pc0"], 32, act=ACT_LEAKY)
        feats = [F_.dense_tc(W["pc1_img"], 32, 32, x1=x, act=ACT_LEAKY) if _tc() else F_.linear(x, W["pc1"], 32, act=ACT_LEAKY)]
        fps = []
        for name, npoint, nsample in self.SA:
            nx, nf, fi = set_abstraction_pm(W[name], npoint, nsample, xyz[-1], feats[-1])
            xyz.append(nx), feats.append(nf), fps.append(fi)

        def h1(t):
            return t[:B]

        def h2(t):
            return t[B:]

        up = set_upconv_pm(W["su3"], 16, xyz[3], xyz[4], feats[3], feats[4])
        cf, cb, ff, flow = cost_volume_pm(W["flow3_r"], h1(xyz[3]), h2(xyz[3]), None, h1(up), None, h2(up), None)
        flows = [flow]
        for lvl, su, fr, dc, k_up, k_warp in ((2, "su2", "flow2_r", "deconv3_2", 5, 5), (1, "su1", "flow1_r", "deconv2_1", 5, 7),
                                             (0, "su0", "flow0_r", "deconv1_0", 7, 7)):
            # the 16-NN lists of the up-conv (dense level in sparse level, both clouds) also serve the four UpsampleFlow
            # calls of cloud 1: their k = 5/7 and k = 3 lists are prefixes (ASF/utils/soflow.py:1459-1461)
            up, idx_up = set_upconv_pm(W[su], 16, xyz[lvl], xyz[lvl + 1], feats[lvl], up, return_idx=True)
            idx_up = idx_up[:B]
            p1, p1s = h1(xyz[lvl]), h1(xyz[lvl + 1])
            p2 = h2(xyz[lvl])
            coarse = upsample_pm(p1, p1s, flow, k_up, idx=idx_up)
            sf_feat = upsample_pm(p1, p1s, ff, k_up, idx=idx_up)
            dcin, dcout = W[dc].shape
            if _tc():
                cfu = F_.dense_tc(W[dc + "_img"], dcout, dcin, x1=upsample_pm(p1, p1s, cf, 3, idx=idx_up), act=ACT_LEAKY)
                cbu = F_.dense_tc(W[dc + "_img"], dcout, dcin, x1=upsample_pm(p1, p1s, cb, 3, idx=idx_up), act=ACT_LEAKY)
            else:
                cfu = F_.linear(upsample_pm(p1, p1s, cf, 3, idx=idx_up), W[dc], dcout, act=ACT_LEAKY)
                cbu = F_.linear(upsample_pm(p1, p1s, cb, 3, idx=idx_up), W[dc], dcout, act=ACT_LEAKY)
            warped = point_warping_pm(p1, p2, coarse, k_warp)
            cf, cb, ff, flow = cost_volume_pm(W[fr], p1, p2, warped, h1(up), cfu, h2(up), cbu, sf=coarse, sf_feat=sf_feat)
            flows.append(flow)
        F_.knn_cache_clear()
        return flows[::-1], [f[:B] for f in fps[:3]]

    @torch.no_grad()
    def forward(self, pc1, pc2, feats1=None, feats2=None):
        """Reference signature: pc1, pc2 f32 [B,3,N] CUDA (, feats1, feats2 f32 [B,C,N]) -> ([4 flows B x3xn], [3 fps idx]).
        As in the reference, the features are used only when both are given (TFlowV3_Occlussion.py:111-116)."""
        nat.require_device()
        x1 = F_.transpose(pc1.contiguous().float())
        x2 = F_.transpose(pc2.contiguous().float())
        if feats1 is None or feats2 is None:
            feats1 = feats2 = None
        else:
            feats1, feats2 = F_.transpose(feats1.contiguous().float()), F_.transpose(feats2.contiguous().float())
        flows, fps = self.forward_pm(x1, x2, feats1, feats2)
        return [F_.transpose(f) for f in flows], fps
