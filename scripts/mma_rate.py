"""Developer tool: measured pacing of tcgen05 kind::tf32 MMAs on this GPU (cycles per M128 x N x K8 instruction)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssf_slam_b200 import _native as nat
nat.require_device()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for N in (32, 64, 128, 256):
    for mode in (0, 1, 2, 3):
        for bufs in (1, 2):
            if bufs * N > 480:
                continue
            reps = 2000
            nat.check(nat.dev_lib().ssf_tc_mma_rate(N, mode, reps, bufs, nat.ptr(out), nat.stream()), nat.dev_lib())
            torch.cuda.synchronize()
            t = out.cpu().tolist()
            print("N=%3d A-from-%s acc_bufs=%d: %.1f cycles/MMA total, %.1f issue" % (N, ("TMEM", "smem", "TMEM/uniform-issue", "smem/uniform-issue")[mode], bufs, t[0] / reps, t[1] / reps))
