"""Isolated, realistically shaped launches of the two tensor-core kernels for `ncu --profile-from-start off`:
the level-0 cost volume (m = 64, N1 = N2 = 8192) and the su0 gather -> GEMM -> max layer (8192 <- 2048, K = N = 64).
Neighbour lists come from real kNN on synthetic CARLA-shaped clouds so the gathers have the hot path's locality."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ssf_slam_b200 import functional as F_, synth
from ssf_slam_b200.model import prepare_weights
from ssf_slam_b200.weights import random_init_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--which", default="cv,dense")
args = ap.parse_args()
B = args.batch
pool = synth.make_sequence(1000, B, 8192)
x1 = torch.from_numpy(np.stack([it["pos1"] for it in pool])).cuda()
x2 = torch.from_numpy(np.stack([it["pos2"] for it in pool])).cuda()
W = prepare_weights(random_init_state_dict(0), torch.device("cuda:0"))
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device="cuda")
idx = F_.knn_idx(16, x1, x2)
idxw = F_.knn_idx(16, x2, x1)
sub = F_.gather_rows(x1, F_.fps(x1, 2048))
idx_up = F_.knn_idx(16, x1, sub)
w0, su0 = W["flow0_r"], W["su0"]
Gab, Hab, H3 = r(B, 8192, 128) * 0.7, r(B, 8192, 128) * 0.7, r(B, 8192, 64) * 0.5
G = r(B, 2048, 64)


def run():
    if "cv" in args.which:
        F_.cost_volume(Gab, Hab, w0, H3, x1, x2, idx, idxw, 64)
    if "dense" in args.which:
        F_.dense_tc(su0["W2_img"], 64, 64, G=G, b1=su0["b1"], Wd1=su0["Wd"], act1=1, idx=idx_up, pos_src=sub, pos_q=x1,
                    bias=su0["b2"], act=1, epi=F_.EPI_MAX)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run()
e1.record()
torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1))
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
