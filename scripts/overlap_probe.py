"""Do the searches of one stream run under the persistent tensor-core kernels of another?  Times the level-0 cost volume and a
8192 x 8192 k = 16 search (B clouds each) alone and launched together on two streams."""
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
ap.add_argument("--batch", type=int, default=32)
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
w0 = W["flow0_r"]
Gab, Hab, H3 = r(B, 8192, 128) * 0.7, r(B, 8192, 128) * 0.7, r(B, 8192, 64) * 0.5
su0 = W["su0"]
sub = F_.gather_rows(x1, F_.fps(x1, 2048))
idx_up = F_.knn_idx(16, x1, sub)
G = r(B, 2048, 64)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def cv():
    F_.cost_volume(Gab, Hab, w0, H3, x1, x2, idx, idxw, 64)


def dense():
    F_.dense_tc(su0["W2_img"], 64, 64, G=G, b1=su0["b1"], Wd1=su0["Wd"], act1=1, idx=idx_up, pos_src=sub, pos_q=x1, bias=su0["b2"], act=1,
                epi=F_.EPI_MAX)


def knn():
    F_.knn_idx(16, x1, x2)


def fps():
    F_.fps(x1, 2048)


def timed(fa, fb, reps=5):
    for _ in range(2):
        if fa: fa()
        if fb: fb()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sa.wait_event(e0), sb.wait_event(e0)
    for _ in range(reps):
        if fa:
            with torch.cuda.stream(sa):
                fa()
        if fb:
            with torch.cuda.stream(sb):
                fb()
    torch.cuda.current_stream().wait_stream(sa)
    torch.cuda.current_stream().wait_stream(sb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


F_.knn_cache_clear()
for na, fa in (("cost_volume", cv), ("dense_tc su0", dense)):
    for nb, fb in (("knn 8192x8192", knn), ("fps 8192->2048", fps)):
        ta, tb, tab = timed(fa, None), timed(None, fb), timed(fa, fb)
        print("%-14s %.3f ms | %-15s %.3f ms | together %.3f ms (sum %.3f, max %.3f)" % (na, ta, nb, tb, tab, ta + tb, max(ta, tb)))
