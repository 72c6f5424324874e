"""CPU: the oracle against the committed golden vectors (which were produced by / checked against the
unmodified reference in the build container, see oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import frontend, point_ops, tflow_port


def test_point_ops_c_and_torch_match_golden(golden_dir, oracle_c):
    g = np.load(os.path.join(golden_dir, "point_ops.npz"))
    xyz, query = g["xyz"], g["query"]
    assert np.array_equal(point_ops.c_fps(xyz, 256), g["fps256"])
    assert np.array_equal(point_ops.furthest_point_sample_torch(torch.from_numpy(xyz[:, :400]), 32).numpy(),
                          point_ops.c_fps(xyz[:, :400], 32))
    for name, k in (("knn16", 16), ("knn7", 7), ("nn3", 3)):
        d, i = point_ops.c_knn(k, query, xyz)
        assert np.array_equal(i, g[name + "_idx"]) and np.array_equal(d, g[name + "_dist"])
        d2, i2 = point_ops.knn_torch(k, torch.from_numpy(query), torch.from_numpy(xyz))
        assert np.array_equal(i2.numpy(), i) and np.array_equal(d2.numpy(), d)
    for r in (0.5, 2.0, 4.0):
        bi, bc = point_ops.c_ball_query(r, 16, xyz, query)
        assert np.array_equal(bi, g["ball_r%g_idx" % r]) and np.array_equal(bc, g["ball_r%g_cnt" % r])
        ti, tc = point_ops.ball_query(r, 16, torch.from_numpy(xyz), torch.from_numpy(query))
        assert np.array_equal(ti.numpy(), bi) and np.array_equal(tc.numpy(), bc)


def test_point_ops_match_reference_twins_golden(golden_dir, oracle_c):
    """a10 pin: tests/golden/point_twins.npz holds the outputs of the reference's own pure-torch twins of the extension
    (farthest_point_sample / knn_point, ASF/utils/utils.py:68-108; query_ball_point, ASF/SetCover.py:39-63), executed
    unmodified by oracle/gen_golden_point_twins.py on tie-free clouds.  Both oracle restatements must reproduce them."""
    g = np.load(os.path.join(golden_dir, "point_twins.npz"))
    assert np.array_equal(point_ops.c_fps(g["fps_xyz"], 512), g["fps_idx"])
    assert np.array_equal(point_ops.furthest_point_sample_torch(torch.from_numpy(g["fps_xyz"]), 512).numpy(), g["fps_idx"])
    q, r = g["knn_query"], g["knn_ref"]
    for k in (3, 8, 16):
        d, i = point_ops.c_knn(k, q, r)
        assert np.array_equal(i, g["knn%d_idx" % k]) and np.allclose(d, g["knn%d_dist" % k], rtol=1e-5, atol=1e-6)
        d2, i2 = point_ops.knn_torch(k, torch.from_numpy(q), torch.from_numpy(r))
        assert np.array_equal(i2.numpy(), g["knn%d_idx" % k])
    for rad in (0.5, 1.0, 2.0, 4.0):
        bi, bc = point_ops.c_ball_query(rad, 16, g["ball_xyz"], g["ball_new_xyz"])
        assert np.array_equal(bi, g["ball_r%g_idx" % rad]) and np.array_equal(bc, g["ball_r%g_cnt" % rad])
        ti, tc = point_ops.ball_query(rad, 16, torch.from_numpy(g["ball_xyz"]), torch.from_numpy(g["ball_new_xyz"]))
        assert np.array_equal(ti.numpy(), bi) and np.array_equal(tc.numpy(), bc)


def test_knn_tie_break_lowest_index():
    xyz = np.zeros((1, 40, 3), np.float32)  # every distance ties
    d, i = point_ops.knn_torch(5, torch.zeros(1, 3, 3), torch.from_numpy(xyz))
    assert np.array_equal(i.numpy()[0, 0], np.arange(5))
    assert np.array_equal(point_ops.furthest_point_sample_torch(torch.from_numpy(xyz), 4).numpy()[0], [0, 0, 0, 0])


def test_ball_query_edge_cases():
    xyz = torch.tensor([[[0.0, 0, 0], [1, 0, 0], [2, 0, 0]]])
    idx, cnt = point_ops.ball_query(1.0, 4, xyz, torch.tensor([[[0.0, 0, 0], [50.0, 0, 0]]]))
    assert idx.tolist() == [[[0, 1, 0, 0], [0, 0, 0, 0]]] and cnt.tolist() == [[2, 0]]  # d <= r*r inclusive; empty -> zeros


def test_scatter_matches_dense_definition():
    g = torch.Generator().manual_seed(0)
    src = torch.randn(2, 50, 3, generator=g)
    index = torch.randint(0, 7, (2, 50), generator=g)
    index[:, 0] = 6
    sm = point_ops.scatter_softmax(src, index, dim=1)
    ss = point_ops.scatter_sum(src, index, dim=1)
    for b in range(2):
        for j in range(7):
            rows = (index[b] == j).nonzero().flatten()
            if len(rows):
                assert torch.allclose(sm[b, rows], torch.softmax(src[b, rows], 0), atol=1e-6)
                assert torch.allclose(ss[b, j], src[b, rows].sum(0), atol=1e-5)


def test_solve_rt_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "solve_rt.npz"))
    R, t = frontend.solve_rt_svd(g["src"], g["dst"])
    assert np.allclose(R, g["R"], atol=1e-12) and np.allclose(t, g["t"], atol=1e-10)
    _, R2, t2 = frontend.pose_from_sums(frontend.kabsch_sums(g["src"].astype(np.float32), g["dst"].astype(np.float32),
                                                               np.ones(len(g["src"]))))
    assert np.allclose(R2, g["R"], atol=1e-6) and np.allclose(t2, g["t"].ravel(), atol=1e-5)  # Horn == SVD Kabsch


def test_quaternion_roundtrip():
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        q2 = frontend.rotation_to_quaternion(R)
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-12


def test_masker_spec_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "masker.npz"))
    a = frontend.masker_spec(g["pos1"], g["flow"], 0.10)
    assert np.array_equal(a["mask"], g["mask_noseg"]) and np.array_equal(a["odom"], g["odom_noseg"])
    assert np.array_equal(a["mask"], g["s_fg_mask"])  # recovers the generator's moving-vehicle points
    b = frontend.masker_spec(g["pos1"], g["flow"], 0.10, sem=g["sem"], inst=g["inst"], movable=(2,))
    assert np.array_equal(b["mask"], g["mask_seg"]) and np.array_equal(b["odom"], g["odom_seg"])
    R, t = frontend.reference_pose(g["pos1"], g["flow"], a["bg_index"])  # the reference's own pose on the same static set
    assert np.allclose(a["R"], R, atol=1e-5) and np.allclose(a["t"], t.ravel(), atol=1e-4)


def test_masker_degenerate_inputs():
    p = np.zeros((2, 3), np.float32)
    out = frontend.masker_spec(p, p, 0.1)
    assert np.allclose(out["R"], np.eye(3)) and out["mask"].tolist() == [0, 0]


@pytest.mark.parametrize("n,name,ch", [(2048, "tflow_n2048.npz", 3), (2048, "tflow_seg4_n2048.npz", 4)])
def test_tflow_port_matches_reference_golden(golden_dir, oracle_c, n, name, ch):
    """3-channel (shipped) and 4-channel (reference flag add_Seg_after_FLow = True) goldens, both written by
    the unmodified reference."""
    g = np.load(os.path.join(golden_dir, name))
    sd = tflow_port.random_init_state_dict(int(g["weight_seed"]), ch)
    pc1 = torch.from_numpy(g["pos1"].T.copy()).unsqueeze(0)
    pc2 = torch.from_numpy(g["pos2"].T.copy()).unsqueeze(0)
    flows, fps = tflow_port.tflow_forward(sd, pc1, pc2)
    for i in range(3):
        assert np.array_equal(fps[i][0].numpy(), g["fps%d" % (i + 1)])
    for i in range(4):
        # bit-exact in the build container; allow for a different CPU's MKL/oneDNN code path elsewhere
        assert flows[i].shape[1] == ch and np.abs(flows[i][0].numpy() - g["flow%d" % i]).max() <= 2e-5


def test_gmm_spec_matches_sklearn_golden_and_live(golden_dir):
    """oracle/gmm.py (restatement of scikit-learn's EM, the reference's noSeg masker) against labels written by scikit-learn
    itself, and against scikit-learn run now from the same initial parameters."""
    from oracle import gmm
    g = np.load(os.path.join(golden_dir, "gmm_mask.npz"))
    for name in ("gt0", "gt3", "gt7", "tiny"):
        spec = gmm.gmm_spec(g[name + "_points"], g[name + "_flow"])
        assert np.array_equal(spec["labels"], g[name + "_sklearn_labels"].astype(np.int64)), name
        assert np.array_equal(spec["mask"], g[name + "_mask"]) and spec["n_iter"] == int(g[name + "_n_iter"]), name
        assert abs(spec["lower_bound"] - float(g[name + "_sklearn_lower_bound"])) < 1e-9
    spec, labels, gm = gmm.check_against_sklearn(g["gt0_points"], g["gt0_flow"])
    assert np.array_equal(labels, spec["labels"]) and gm.n_iter_ == spec["n_iter"]
    assert np.array_equal(gmm.reference_bg_index(labels), spec["bg_index"])


def test_tflow_port_afterpc_variant_matches_reference_golden(golden_dir, oracle_c):
    """4-channel INPUT variant (TFlowV3_Occlussion_addSeg_afterPC.py:68); golden written by that unmodified file."""
    g = np.load(os.path.join(golden_dir, "tflow_afterpc_n2048.npz"))
    sd = tflow_port.random_init_state_dict(int(g["weight_seed"]), 3, input_channels=4)
    pc1 = torch.from_numpy(g["pos1"].T.copy()).unsqueeze(0)
    pc2 = torch.from_numpy(g["pos2"].T.copy()).unsqueeze(0)
    f1 = torch.cat([pc1, torch.from_numpy(g["lab1"])[None, None]], dim=1)
    f2 = torch.cat([pc2, torch.from_numpy(g["lab2"])[None, None]], dim=1)
    flows, fps = tflow_port.tflow_forward(sd, pc1, pc2, feats1=f1, feats2=f2)
    for i in range(3):
        assert np.array_equal(fps[i][0].numpy(), g["fps%d" % (i + 1)])
    for i in range(4):
        assert np.abs(flows[i][0].numpy() - g["flow%d" % i]).max() <= 2e-5
