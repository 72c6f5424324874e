"""oracle/tflow_port.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional CPU restatement (plain torch fp32) of the reference's ActiveSceneFlow ``TFlow``
forward pass, written against a reference-format ``state_dict`` (317 tensors, same key names
as ``scripts/ActiveSceneFlow/TFlowV3_Occlussion.py`` produces).  It exists because
``/root/reference`` cannot travel to the GPU box: this port is the checker the parity tests
and ``smoke()`` compare the CUDA path against, and the CPU baseline ``bench.py`` times.

Pinned: oracle/gen_golden.py runs the UNMODIFIED reference model (imported from
/root/reference over the oracle shims) and this port on the same seeded weights and clouds and
asserts bit-identical flows and FPS indices; the resulting vectors are committed under
tests/golden/.  The point operators underneath are the oracle's (oracle/point_ops.py; the
extension itself is absent from the reference; pinned against the reference's in-tree pure-torch
twins on tie-free data, tie-breaking by written specification -- see oracle/point_ops.py).

Reference lines followed, per function, are cited in each docstring.  Layout is the
reference's: features ``[B,C,N]``, coordinates ``[B,3,N]``.
"""
import torch
import torch.nn.functional as F

from oracle import point_ops as ops

LEAKY = 0.1  # ASF/utils/soflow.py:17, ASF/TFlowV3_Occlussion.py:18


def _t(x):
    return x.permute(0, 2, 1).contiguous()


def _bn(sd, key, x):
    """Eval-mode BatchNorm with running statistics (nn.BatchNorm{1,2}d defaults, eps 1e-5)."""
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"],
                        sd[key + ".weight"], sd[key + ".bias"], False, 0.1, 1e-5)


def _conv(sd, key, x):
    w = sd[key + ".weight"]
    b = sd.get(key + ".bias")
    return F.conv2d(x, w, b) if w.dim() == 4 else F.conv1d(x, w, b)


def _gather_rows(points_bnc, idx):
    """index_points_group, ASF/utils/soflow.py:21-32: [B,N,C] x [B,M,K] -> [B,M,K,C]."""
    return ops.grouping_operation(_t(points_bnc).float(), idx.int()).permute(0, 2, 3, 1).contiguous()


def leaky_conv1d(sd, prefix, x):
    """Conv1d blocks: model-local one has no bias (TFlowV3_Occlussion.py:22-38), soflow's has
    (soflow.py:1260-1276); both k=1 + LeakyReLU(0.1), no BN (use_bn=False)."""
    return F.leaky_relu(_conv(sd, prefix + ".composed_module.0", x), LEAKY)


def set_abstraction(sd, prefix, npoint, nsample, xyz, feats):
    """PointNetSetAbstraction.forward, ASF/utils/utils.py:208-248 (kNN grouping, radius unused)."""
    xyz_t = _t(xyz)
    fps_idx = ops.furthest_point_sample(xyz_t, npoint)
    new_xyz = ops.gather_operation(xyz, fps_idx)
    _, knn_idx = ops.knn(nsample, _t(new_xyz), xyz_t)
    pos_diff = ops.grouping_operation(xyz, knn_idx) - new_xyz.unsqueeze(-1)
    x = torch.cat([pos_diff, ops.grouping_operation(feats, knn_idx)], dim=1)
    n_layers = len([k for k in sd if k.startswith(prefix + ".mlp_convs.") and k.endswith(".weight")])
    for i in range(n_layers):
        x = F.relu(_bn(sd, "%s.mlp_bns.%d" % (prefix, i), _conv(sd, "%s.mlp_convs.%d" % (prefix, i), x)))
    return new_xyz, torch.max(x, -1)[0], fps_idx


def set_upconv(sd, prefix, nsample, pos1, pos2, feat1, feat2):
    """PointNetSetUpConv.forward, ASF/utils/utils.py:274-315 (knn=True, sf1=sf2=None)."""
    B, _, N = pos1.shape
    _, idx = ops.knn(nsample, _t(pos1), _t(pos2))
    pos_diff = ops.grouping_operation(pos2.float(), idx) - pos1.view(B, -1, N, 1)
    x = torch.cat([ops.grouping_operation(feat2.float(), idx), pos_diff], dim=1)
    i = 0
    while "%s.mlp1_convs.%d.0.weight" % (prefix, i) in sd:
        p = "%s.mlp1_convs.%d" % (prefix, i)
        x = F.relu(_bn(sd, p + ".1", _conv(sd, p + ".0", x)))
        i += 1
    x = x.max(-1)[0]
    if feat1 is not None:
        x = torch.cat([x, feat1], dim=1)
    i = 0
    while "%s.mlp2_convs.%d.0.weight" % (prefix, i) in sd:
        p = "%s.mlp2_convs.%d" % (prefix, i)
        x = F.relu(_bn(sd, p + ".1", _conv(sd, p + ".0", x)))
        i += 1
    return x


def _inv_dist_interp(query, src_pos, src_val, idx):
    """Normalised inverse-distance interpolation shared by UpsampleFlow / PointWarping
    (soflow.py:1462-1471, :1244-1250): weights from gathered positions, distance clamp 1e-10."""
    B, C, N = query.shape
    k = idx.shape[-1]
    rel = ops.grouping_operation(src_pos, idx) - query.view(B, C, N, 1)
    dist = torch.norm(rel, dim=1).clamp(min=1e-10)
    norm = torch.sum(1.0 / dist, dim=2, keepdim=True)
    weight = (1.0 / dist) / norm
    return torch.sum(weight.view(B, 1, N, k) * ops.grouping_operation(src_val.float(), idx), dim=-1)


def upsample_flow(xyz, sparse_xyz, sparse_flow, k=3):
    """UpsampleFlow.forward, ASF/utils/soflow.py:1443-1475."""
    if k == 3:
        _, idx = ops.three_nn(_t(xyz), _t(sparse_xyz))
    else:
        _, idx = ops.knn(k, _t(xyz), _t(sparse_xyz))
    return _inv_dist_interp(xyz, sparse_xyz, sparse_flow, idx).clamp(-100.0, 100.0)


def point_warping(pos1, pos2, flow1, nsample):
    """PointWarping.forward, ASF/utils/soflow.py:1223-1257 (positions clamped to +-10 m)."""
    if flow1 is None:
        return pos2
    moved = pos1 + flow1[:, 0:3, :]   # 4-channel flows (add_Seg_after_FLow, soflow.py:1228) move by their first three
    if nsample is None:
        _, idx = ops.three_nn(_t(pos2), _t(moved))
    else:
        _, idx = ops.knn(nsample, _t(pos2), _t(moved))
    flow2 = _inv_dist_interp(pos2, moved, flow1, idx)
    return (pos2 - flow2[:, 0:3, :]).clamp(-10.0, 10.0)   # soflow.py:1252-1257


def _weightnet(sd, p, x):
    """weightnet1, ASF/utils/soflow.py:307-313: conv(no bias)+BN+ReLU, conv(no bias)+BN+ReLU, conv(bias)."""
    x = F.relu(_bn(sd, p + ".1", _conv(sd, p + ".0", x)))
    x = F.relu(_bn(sd, p + ".4", _conv(sd, p + ".3", x)))
    return _conv(sd, p + ".6", x)


def _mlp(sd, p, x):
    i = 0
    while "%s.%d.weight" % (p, i) in sd:
        x = F.leaky_relu(_conv(sd, "%s.%d" % (p, i), x), LEAKY)
        i += 1
    return x


def cost_volume(sd, prefix, nsample, use_flow, xyz1, xyz2, xyz2w, points1, points2, sf=None, sf_feat=None,
                return_aux=False):
    """PointConvTransFlowV2.forward, ASF/utils/soflow.py:354-525; dataflow = SURVEY.md Appendix F."""
    B, C, N1 = xyz1.shape
    S = nsample
    x1, x2 = _t(xyz1), _t(xyz2)
    x2w = _t(xyz2w) if xyz2w is not None else x2
    p1, p2 = _t(points1), _t(points2)
    D1 = p1.shape[-1]

    if sf is not None and use_flow:
        sf_t = _t(sf)
        _, idx = ops.knn(S, x1 + sf_t[:, :, 0:3], x2)   # soflow.py:386-389 (identity slice for 3-channel flows)
    else:
        if sf is not None:
            sf_t = _t(sf)
        _, idx = ops.knn(S, x1, x2)
    direction = _gather_rows(x2, idx) - x1.view(B, N1, 1, C)
    g1 = p1.view(B, N1, 1, D1).repeat(1, 1, S, 1)
    a = torch.cat([g1, _gather_rows(p2, idx)], dim=-1).permute(0, 3, 2, 1)
    a = _mlp(sd, prefix + ".mlp_convs", a)

    _, idxw = ops.knn(S, x1, x2w)
    directionw = _gather_rows(x2, idxw) - x1.view(B, N1, 1, C)
    aw = torch.cat([g1, _gather_rows(p2, idxw)], dim=-1).permute(0, 3, 2, 1)
    aw = _mlp(sd, prefix + ".mlp_convs2", aw)

    qk = torch.matmul(a.permute(0, 3, 2, 1).contiguous(), aw.permute(0, 3, 1, 2).contiguous())
    qk = torch.softmax(qk, -2) * torch.softmax(qk, -1)

    dir_c = direction.permute(0, 3, 2, 1).contiguous()
    dirw_c = directionw.permute(0, 3, 2, 1).contiguous()
    if sf_feat is not None:
        sff = _t(sf_feat)
        gsf = sff.view(B, N1, 1, sff.shape[-1]).repeat(1, 1, S, 1).permute(0, 3, 2, 1).contiguous()
        c = torch.cat([a, gsf, dir_c], dim=1)
        cw = torch.cat([aw, gsf, dirw_c], dim=1)
    else:
        c = torch.cat([a, dir_c], dim=1)
        cw = torch.cat([aw, dirw_c], dim=1)
    c = _mlp(sd, prefix + ".mlp_convs3", c)
    cw = _mlp(sd, prefix + ".mlp_convs3", cw)

    a_mix = a + torch.matmul(qk, aw.permute(0, 3, 2, 1).contiguous()).permute(0, 3, 2, 1).contiguous()
    aw_mix = aw + torch.matmul(a.permute(0, 3, 1, 2).contiguous(), qk).permute(0, 2, 3, 1).contiguous()
    g = _weightnet(sd, prefix + ".weightnet1", a_mix)
    gw = _weightnet(sd, prefix + ".weightnet1", aw_mix)
    w_fwd = torch.softmax(g, dim=2)

    key = idxw.view(B, -1).long()
    cw_flat = cw.permute(0, 3, 2, 1).reshape(B, -1, c.shape[1])
    w_bwd = ops.scatter_softmax(gw.permute(0, 3, 2, 1).reshape(B, -1, gw.shape[1]), key, dim=1)
    # torch_scatter sizes the output as key.max()+1 (soflow.py:481); the reference then gathers it with knn_idx
    # (:489) and upsamples it with indices up to N2-1, a latent out-of-bounds read whenever the highest-index pc2
    # point is nobody's neighbour.  The oracle allocates all N2 rows (zeros for
    # unreferenced points): identical wherever the reference is defined.
    cost_bwd = ops.scatter_sum(cw_flat * w_bwd, key, dim=1, dim_size=xyz2.shape[2])
    cost_fwd = torch.sum(w_fwd * c, dim=2)

    g_bwd = _gather_rows(cost_bwd, idx)
    # memory reinterpretation, not a transpose (soflow.py:490)
    g_fwd = cost_fwd.view(B, N1, 1, cost_fwd.shape[1]).repeat(1, 1, S, 1)
    if sf_feat is not None:
        x = torch.cat([g_fwd, g_bwd, gsf.permute(0, 3, 2, 1).contiguous(), direction], dim=-1)
    else:
        x = torch.cat([g_fwd, g_bwd, direction], dim=-1)
    x = _mlp(sd, prefix + ".mlp_convs4", x.permute(0, 3, 2, 1))
    feats = torch.max(x, dim=2)[0]
    i = 0
    while "%s.flow_mlp_convs.%d.composed_module.0.weight" % (prefix, i) in sd:
        feats = leaky_conv1d(sd, "%s.flow_mlp_convs.%d" % (prefix, i), feats)
        i += 1
    re_sf = _conv(sd, prefix + ".fc", feats).clamp(-50.0, 50.0)
    if sf is not None:
        re_sf = re_sf + sf
    out = (cost_fwd, cost_bwd.permute(0, 2, 1).contiguous(), feats, re_sf.clamp(-50.0, 50.0))
    if return_aux:
        return out, {"idx": idx, "idxw": idxw, "a": a, "aw": aw, "c": c, "cw": cw, "g": g, "gw": gw}
    return out


def refine_flow(sd, prefix, nsample, use_flow, pc1, pc2, feats1, feats2, warp_k, c_flow=None, flow_feats=None):
    """RefineFlowRegressor.forward, ASF/TFlowV3_Occlussion.py:51-62."""
    pc2_warp = None if c_flow is None else point_warping(pc1, pc2, c_flow, warp_k)
    if use_flow:
        return cost_volume(sd, prefix + ".cost", nsample, True, pc1, pc2, pc2_warp, feats1, feats2, c_flow, flow_feats)
    return cost_volume(sd, prefix + ".cost", nsample, False, pc1, pc2, pc2_warp, feats1, feats2)


# (prefix, npoint, nsample) per pyramid level, ASF/TFlowV3_Occlussion.py:70-77
SA_CFG = [("sa1", 2048, 16), ("sa2", 512, 16), ("sa3", 256, 16), ("sa4", 128, 8)]


@torch.no_grad()
def tflow_forward(sd, pc1, pc2, return_intermediates=False, feats1=None, feats2=None):
    """TFlow.forward, ASF/TFlowV3_Occlussion.py:105-196.  pc1, pc2: f32 [B,3,N] on CPU.  feats1, feats2 [B,C,N]: the optional
    input features (used only when both are given, :111-116; C = 4 for TFlowV3_Occlussion_addSeg_afterPC.py:68)."""
    sd = {k: v for k, v in sd.items()}
    inter = {}

    def point_conv(x):
        return leaky_conv1d(sd, "point_conv.1", leaky_conv1d(sd, "point_conv.0", x))

    in1, in2 = (pc1, pc2) if feats1 is None or feats2 is None else (feats1, feats2)
    pcs1, pcs2, f1, f2, fps = [pc1], [pc2], [point_conv(in1)], [point_conv(in2)], []
    for prefix, npoint, nsample in SA_CFG:
        x, f, i = set_abstraction(sd, prefix, npoint, nsample, pcs1[-1], f1[-1])
        pcs1.append(x), f1.append(f), fps.append(i)
        x, f, _ = set_abstraction(sd, prefix, npoint, nsample, pcs2[-1], f2[-1])
        pcs2.append(x), f2.append(f)

    # level 3
    u1 = set_upconv(sd, "su3", 16, pcs1[3], pcs1[4], f1[3], f1[4])
    u2 = set_upconv(sd, "su3", 16, pcs2[3], pcs2[4], f2[3], f2[4])
    cf, cb, ff, flow = refine_flow(sd, "flow3_r", 16, False, pcs1[3], pcs2[3], u1, u2, 3)
    flows = [flow]
    inter["l3"] = (cf, cb, ff, flow)
    inter["in3"] = dict(u1=u1, u2=u2)

    # levels 2, 1, 0: (su, flow regressor, deconv, k for flow/feat upsample, warp k)
    for lvl, su, fr, dc, k_up, k_warp in ((2, "su2", "flow2_r", "deconv3_2", 5, 5),
                                         (1, "su1", "flow1_r", "deconv2_1", 5, 7),
                                         (0, "su0", "flow0_r", "deconv1_0", 7, 7)):
        prev1, prev2 = u1, u2
        u1 = set_upconv(sd, su, 16, pcs1[lvl], pcs1[lvl + 1], f1[lvl], u1)
        u2 = set_upconv(sd, su, 16, pcs2[lvl], pcs2[lvl + 1], f2[lvl], u2)
        coarse = upsample_flow(pcs1[lvl], pcs1[lvl + 1], flow, k=k_up)
        sf_feat = upsample_flow(pcs1[lvl], pcs1[lvl + 1], ff, k=k_up)
        cfu = leaky_conv1d(sd, dc, upsample_flow(pcs1[lvl], pcs1[lvl + 1], cf))
        cbu = leaky_conv1d(sd, dc, upsample_flow(pcs1[lvl], pcs1[lvl + 1], cb))
        in1 = torch.cat([u1, cfu], dim=1)
        in2 = torch.cat([u2, cbu], dim=1)
        cf, cb, ff, flow = refine_flow(sd, fr, 16, True, pcs1[lvl], pcs2[lvl], in1, in2, k_warp, coarse, sf_feat)
        flows.append(flow)
        inter["l%d" % lvl] = (cf, cb, ff, flow)
        # what each layer of this level was fed (teacher forcing in the parity tests)
        inter["in%d" % lvl] = dict(prev1=prev1, prev2=prev2, u1=u1, u2=u2, coarse=coarse, sf_feat=sf_feat, cfu=cfu, cbu=cbu,
                                   su=su, fr=fr, dc=dc, k_up=k_up, k_warp=k_warp)

    out = (flows[::-1], fps[:3])
    if return_intermediates:
        inter.update(pcs1=pcs1, pcs2=pcs2, f1=f1, f2=f2, fps=fps)
        return out, inter
    return out


from ssf_slam_b200.weights import random_init_state_dict  # noqa: E402,F401  (deterministic reference-format weights)
