"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Run in the BUILD container only (needs /root/reference).

Pins the oracle and writes the committed fixtures under tests/golden/:

1. tflow_n{N}.npz -- the UNMODIFIED reference ``TFlow`` (imported from /root/reference over the oracle
   shims) is loaded with ``oracle.tflow_port.random_init_state_dict(seed)`` and run on a seeded synthetic
   frame pair; the oracle port must reproduce its four flows bit-for-bit and its FPS indices exactly, or
   this script aborts.  Stored: inputs, reference outputs, seeds.
2. solve_rt.npz -- the reference's own ``slove_RT_by_SVD`` source text (scripts/PointCloudOdometry.py:15-33)
   is exec'd verbatim and compared with oracle.frontend.solve_rt_svd; stored with inputs.
3. point_ops.npz -- small clouds with exact duplicates: FPS / kNN / 3-NN / ball-query results of the C and
   the torch restatements (which must agree).  The extension is absent from the reference, so these pin
   the written tie-breaking spec; the tie-free behaviour is pinned against the reference's own pure-torch twins by
   oracle/gen_golden_point_twins.py (tests/golden/point_twins.npz).
4. masker.npz -- oracle.frontend.masker_spec on a synthetic frame with gt flow (+ noise), noSeg and Seg.

Usage:  python -m oracle.gen_golden
"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import frontend, point_ops, tflow_port  # noqa: E402
from oracle.ref_harness import import_reference_tflow  # noqa: E402
from ssf_slam_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def gen_tflow(n_points, data_seed, weight_seed=0, flow_channels=3, with_labels=False):
    """flow_channels=4: the reference with its source flag ``add_Seg_after_FLow`` (utils/datasets/carla.py:9, imported by
    name into utils/soflow.py:8) switched on for the duration of the call -- 4-channel flow heads."""
    TFlow = import_reference_tflow()
    import utils.soflow as ref_soflow  # the reference module whose global the classes read
    ref_soflow.add_Seg_after_FLow = flow_channels == 4
    try:
        _gen_tflow(TFlow, n_points, data_seed, weight_seed, flow_channels, with_labels)
    finally:
        ref_soflow.add_Seg_after_FLow = False


def _gen_tflow(TFlow, n_points, data_seed, weight_seed, flow_channels, with_labels=False):
    sd = tflow_port.random_init_state_dict(weight_seed, flow_channels)
    net = TFlow().eval()
    net.load_state_dict(sd, strict=True)
    item = synth.make_pair(data_seed, n_points)
    pc1 = torch.from_numpy(item["pos1"].T.copy()).unsqueeze(0)
    pc2 = torch.from_numpy(item["pos2"].T.copy()).unsqueeze(0)
    with torch.no_grad():
        flows, fps = net(pc1, pc2)
    pflows, pfps = tflow_port.tflow_forward(sd, pc1, pc2)
    for a, b in zip(flows, pflows):
        assert torch.equal(a, b), "oracle port deviates from the reference"
    for a, b in zip(fps, pfps):
        assert torch.equal(a, b), "oracle port FPS deviates from the reference"
    name = "tflow_n%d.npz" % n_points if flow_channels == 3 else "tflow_seg4_n%d.npz" % n_points
    extra = dict(sem=item["sem"], inst=item["inst"]) if with_labels else {}   # config 3 (Seg pipeline): synthetic labels
    np.savez_compressed(os.path.join(OUT, name), pos1=item["pos1"], pos2=item["pos2"], **extra,
                        flow0=flows[0][0].numpy(), flow1=flows[1][0].numpy(), flow2=flows[2][0].numpy(),
                        flow3=flows[3][0].numpy(), fps1=fps[0][0].numpy(), fps2=fps[1][0].numpy(), fps3=fps[2][0].numpy(),
                        weight_seed=weight_seed, data_seed=data_seed, flow_channels=flow_channels)
    print("%s: reference == port (bit-exact); |flow|max %.4f" % (name, float(flows[0].abs().max())))


def gen_tflow_afterpc(n_points, data_seed, weight_seed=0):
    """The 4-channel INPUT variant: the unmodified ``TFlowV3_Occlussion_addSeg_afterPC.TFlow`` (first layer Conv1d(4, 32)),
    called as its driver does with ``[xyz | label]`` features."""
    TFlow = import_reference_tflow("TFlowV3_Occlussion_addSeg_afterPC")
    sd = tflow_port.random_init_state_dict(weight_seed, 3, input_channels=4)
    net = TFlow().eval()
    net.load_state_dict(sd, strict=True)
    item = synth.make_pair(data_seed, n_points)
    pc1 = torch.from_numpy(item["pos1"].T.copy()).unsqueeze(0)
    pc2 = torch.from_numpy(item["pos2"].T.copy()).unsqueeze(0)
    lab1 = torch.from_numpy(item["s_fg_mask"].astype(np.float32))[None, None]
    lab2 = torch.from_numpy(item["t_fg_mask"].astype(np.float32))[None, None]
    f1, f2 = torch.cat([pc1, lab1], dim=1), torch.cat([pc2, lab2], dim=1)
    with torch.no_grad():
        flows, fps = net(pc1, pc2, f1, f2)
    pflows, pfps = tflow_port.tflow_forward(sd, pc1, pc2, feats1=f1, feats2=f2)
    for a, b in zip(list(flows) + list(fps), list(pflows) + list(pfps)):
        assert torch.equal(a, b), "oracle port deviates from the reference (afterPC variant)"
    name = "tflow_afterpc_n%d.npz" % n_points
    np.savez_compressed(os.path.join(OUT, name), pos1=item["pos1"], pos2=item["pos2"], lab1=lab1[0, 0].numpy(), lab2=lab2[0, 0].numpy(),
                        flow0=flows[0][0].numpy(), flow1=flows[1][0].numpy(), flow2=flows[2][0].numpy(),
                        flow3=flows[3][0].numpy(), fps1=fps[0][0].numpy(), fps2=fps[1][0].numpy(), fps3=fps[2][0].numpy(),
                        weight_seed=weight_seed, data_seed=data_seed)
    print("%s: reference == port (bit-exact); |flow|max %.4f" % (name, float(flows[0].abs().max())))


def _reference_solve_rt():
    path = "/root/reference/scripts/PointCloudOdometry.py"
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "slove_RT_by_SVD"][0]
    ns = {"np": np}
    exec(compile(ast.Module([fn], []), path, "exec"), ns)
    return ns["slove_RT_by_SVD"]


def gen_solve_rt():
    ref = _reference_solve_rt()
    rng = np.random.default_rng(11)
    src = rng.uniform(-50, 50, (500, 3))
    yaw, pitch = 0.03, -0.01
    Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
    dst = src @ (Rz @ Ry).T + np.array([1.2, -0.1, 0.02]) + 0.01 * rng.standard_normal((500, 3))
    R, t = ref(src, dst)
    R2, t2 = frontend.solve_rt_svd(src, dst)
    assert np.array_equal(R, R2) and np.array_equal(t, t2), "solve_rt_svd deviates from the reference"
    np.savez_compressed(os.path.join(OUT, "solve_rt.npz"), src=src, dst=dst, R=R, t=t)
    print("solve_rt: reference == oracle (bit-exact)")


def gen_point_ops():
    rng = np.random.default_rng(5)
    B, N = 2, 1500
    xyz = rng.uniform(-30, 30, (B, N, 3)).astype(np.float32)
    xyz[:, 300:500] = xyz[:, 0:200]  # exact duplicates -> exact distance ties
    xyz[:, 700:720] = xyz[:, 0:1]    # a 21-fold duplicate (more copies than k)
    query = np.concatenate([xyz[:, ::5], rng.uniform(-30, 30, (B, 100, 3)).astype(np.float32)], 1)
    out = dict(xyz=xyz, query=query)
    saved = point_ops.USE_C
    res = {}
    for use_c in (True, False):
        point_ops.USE_C = use_c
        x, q = torch.from_numpy(xyz), torch.from_numpy(query)
        res[use_c] = dict(
            fps=point_ops.furthest_point_sample(x, 256).numpy(),
            knn16=point_ops.knn(16, q, x), knn7=point_ops.knn(7, q, x), nn3=point_ops.three_nn(q, x))
    point_ops.USE_C = saved
    assert np.array_equal(res[True]["fps"], res[False]["fps"])
    for k in ("knn16", "knn7", "nn3"):
        assert torch.equal(res[True][k][0], res[False][k][0]) and torch.equal(res[True][k][1], res[False][k][1]), k
    out["fps256"] = res[True]["fps"]
    for k in ("knn16", "knn7", "nn3"):
        out[k + "_dist"] = res[True][k][0].numpy()
        out[k + "_idx"] = res[True][k][1].numpy()
    for r in (0.5, 2.0, 4.0):
        bi, bc = point_ops.ball_query(r, 16, torch.from_numpy(xyz), torch.from_numpy(query))
        ci, cc = point_ops.c_ball_query(r, 16, xyz, query)
        assert np.array_equal(bi.numpy(), ci) and np.array_equal(bc.numpy(), cc)
        out["ball_r%g_idx" % r] = ci
        out["ball_r%g_cnt" % r] = cc
    np.savez_compressed(os.path.join(OUT, "point_ops.npz"), **out)
    print("point_ops: C == torch restatement")


def gen_masker():
    item = synth.make_pair(2000, 8192)
    rng = np.random.default_rng(3)
    flow = (item["gt"] + 0.02 * rng.standard_normal(item["gt"].shape)).astype(np.float32)
    a = frontend.masker_spec(item["pos1"], flow, 0.10)
    b = frontend.masker_spec(item["pos1"], flow, 0.10, sem=item["sem"], inst=item["inst"], movable=synth.MOVABLE_CLASSES)
    np.savez_compressed(os.path.join(OUT, "masker.npz"), pos1=item["pos1"], flow=flow, sem=item["sem"], inst=item["inst"],
                        s_fg_mask=item["s_fg_mask"], mask_noseg=a["mask"], odom_noseg=a["odom"], mask_seg=b["mask"],
                        odom_seg=b["odom"], R0_noseg=a["R0"], t0_noseg=a["t0"])
    print("masker: noSeg dyn %d, Seg dyn %d, gt fg %d" % (a["mask"].sum(), b["mask"].sum(), item["s_fg_mask"].sum()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_solve_rt()
    gen_point_ops()
    gen_masker()
    gen_tflow(2048, data_seed=42)
    gen_tflow(8192, data_seed=0)
    gen_tflow(2048, data_seed=43, flow_channels=4)
    gen_tflow_afterpc(2048, data_seed=44)
    gen_tflow(16384, data_seed=3000, with_labels=True)   # BASELINE config 3 size
