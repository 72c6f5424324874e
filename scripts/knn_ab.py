import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ssf_slam_b200 import functional as F_, synth
B = 32
pool = synth.make_sequence(1000, B, 8192)
x1 = torch.from_numpy(np.stack([it["pos1"] for it in pool])).cuda()
x2 = torch.from_numpy(np.stack([it["pos2"] for it in pool])).cuda()
sub = F_.gather_rows(x1, F_.fps(x1, 2048))
def t(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
print(os.environ.get("TAG"), "k16 8192x8192 %.3f  k7 8192x8192 %.3f  k16 8192x2048 %.3f  k16 2048x8192 %.3f  k16 2048x2048 %.3f" % (
    t(lambda: F_.knn_idx(16, x1, x2)), t(lambda: F_.knn_idx(7, x1, x2)), t(lambda: F_.knn_idx(16, x1, sub)), t(lambda: F_.knn_idx(16, sub, x1)), t(lambda: F_.knn_idx(16, sub, sub))))


def build(ref):
    F_.knn_cache_clear()
    F_._knn_blocks(ref)


print(os.environ.get("TAG"), "index build (B=%d): 8192 pts %.3f ms  2048 pts %.3f ms" % (B, t(lambda: build(x2)), t(lambda: build(sub))))
