"""Timings of furthest point sampling at the shapes of the step (A/B: SSF_FPS_PRUNE=0 selects the plain kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ssf_slam_b200 import functional as F_, synth

B = int(os.environ.get("B", "128"))
pool = synth.make_sequence(1000, min(B, 64), 8192)
x1 = torch.from_numpy(np.stack([it["pos1"] for it in pool])).cuda()
x1 = x1.repeat((B + x1.shape[0] - 1) // x1.shape[0], 1, 1)[:B].contiguous()


def t(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5


l1 = F_.gather_rows(x1, F_.fps(x1, 2048))
l2 = F_.gather_rows(l1, F_.fps(l1, 512))
l3 = F_.gather_rows(l2, F_.fps(l2, 256))
print(os.environ.get("SSF_FPS_PRUNE", "1"), "B=%d  8192->2048 %.3f ms  2048->512 %.3f ms  512->256 %.3f ms  4096->1024 %.3f ms" % (
    B, t(lambda: F_.fps(x1, 2048)), t(lambda: F_.fps(l1, 512)), t(lambda: F_.fps(l2, 256)), t(lambda: F_.fps(x1[:, :4096].contiguous(), 1024))))
