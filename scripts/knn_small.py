import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ssf_slam_b200 import functional as F_, synth
B = 128
pool = synth.make_sequence(1000, 64, 8192)
x = torch.from_numpy(np.stack([it["pos1"] for it in pool])).cuda().repeat(2,1,1)
l1 = F_.gather_rows(x, F_.fps(x, 2048)); l2 = F_.gather_rows(l1, F_.fps(l1, 512)); l3 = F_.gather_rows(l2, F_.fps(l2, 256)); l4 = F_.gather_rows(l3, F_.fps(l3, 128))
def t(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for mn in (512, 64):
    F_.KNN_BLOCKS_MIN_REF = mn
    def run(k, q, r):
        F_.knn_cache_clear()
        return F_.knn_idx(k, q, r)
    print("min_ref", mn, " ".join("%s %.3f" % (n, t(lambda: run(k, q, r))) for n, k, q, r in [
        ("k16 256x256", 16, l3, l3), ("k16 512x256", 16, l2, l3), ("k16 256x128", 16, l3, l4), ("k8 128x256", 8, l4, l3), ("k16 256x512", 16, l3, l2), ("k16 512x512",16,l2,l2)]))
a = {}
for mn in (512, 64):
    F_.KNN_BLOCKS_MIN_REF = mn; F_.knn_cache_clear()
    a[mn] = [F_.knn_idx(16, l2, l3), F_.knn_idx(16, l3, l4), F_.knn_idx(8, l4, l3)]
print("identical:", all(torch.equal(u, v) for u, v in zip(a[512], a[64])))
