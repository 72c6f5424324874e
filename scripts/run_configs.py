"""BASELINE.json configs 2 (index 2: Seg_ActiveSceneFlow, per-instance voting, N = 16384) and 4 (index 4: dense full-beam
stress, N = 65536, FPS / kNN / ball-query radius sweep) on one B200: measured numbers for profiles/, one JSON object on stdout.
The bench line (bench.py) is configs[1]; these are the other single-GPU configurations, timed the same way (CUDA events on the
launching stream, 3 warm-ups)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ssf_slam_b200 import functional as F_, pointnet2_utils as pu, synth
from ssf_slam_b200.model import TFlow
from ssf_slam_b200.weights import random_init_state_dict


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
# ---- config index 2: Seg pipeline, synthetic semantic / instance labels, N = 16384
N, B = 16384, 16
pool = synth.make_sequence(2000, B, N)
x1 = torch.from_numpy(np.stack([it["pos1"] for it in pool])).cuda()
x2 = torch.from_numpy(np.stack([it["pos2"] for it in pool])).cuda()
sem = torch.from_numpy(np.stack([it["sem"] for it in pool]).astype(np.int32)).cuda()
inst = torch.from_numpy(np.stack([it["inst"] for it in pool]).astype(np.int32)).cuda()
n_inst = int(inst.max().item()) + 1
net = TFlow()
net.load_state_dict(random_init_state_dict(0))
net = net.cuda().eval()


def seg_step():
    flows, _ = net.forward_pm(x1, x2)
    return F_.frontend(x1, flows[0], mode=1, sem=sem, movable=synth.MOVABLE_CLASSES, inst=inst, n_inst=n_inst, tau=0.10)


ms = timed(seg_step, reps=4)
out["config2_seg_n16384"] = {"pairs_per_step": B, "ms_per_step": ms, "frame_pairs_per_s": B / ms * 1e3, "n_instances": n_inst,
                             "note": "flow + residual masker with semantic seed and per-instance voting + pose, device resident"}
del net, x1, x2, sem, inst

# ---- config index 4: N = 65536 uniform-density full-beam clouds
N, B = 65536, 4
g = torch.Generator(device="cuda").manual_seed(5000)
xyz = (torch.rand(B, N, 3, device="cuda", generator=g) - 0.5) * torch.tensor([200.0, 200.0, 20.0], device="cuda")
ms = timed(lambda: pu.furthest_point_sample(xyz, 2048), reps=3, warm=1)
fps_idx = pu.furthest_point_sample(xyz, 2048)
new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
stress = {"clouds": B, "fps_npoint2048": {"ms": ms, "G_point_updates_per_s": B * N * 2048.0 / ms / 1e6}}
ms = timed(lambda: pu.knn(16, xyz, xyz), reps=2, warm=1)
stress["knn_k16_65536x65536"] = {"ms": ms, "G_pair_evals_per_s": B * float(N) * N / ms / 1e6, "bytes_algorithmic": B * (24.0 * N + 8.0 * N * 16)}
ms = timed(lambda: pu.knn(16, new_xyz, xyz), reps=3, warm=1)
stress["knn_k16_2048x65536"] = {"ms": ms, "G_pair_evals_per_s": B * 2048.0 * N / ms / 1e6}
for ns in (16, 32):
    for r in (0.5, 1.0, 2.0, 4.0):
        ms = timed(lambda: pu.ball_query(r, ns, xyz, new_xyz), reps=3, warm=1)
        _, cnt = pu.ball_query(r, ns, xyz, new_xyz, return_count=True)
        stress["ball_query_r%g_ns%d" % (r, ns)] = {"ms": ms, "G_pair_evals_per_s": B * 2048.0 * N / ms / 1e6,
                                                  "mean_in_range": float(cnt.float().mean().item())}
out["config4_stress_n65536"] = stress
print(json.dumps(out))
