"""Developer experiment: a 256-column plain-row layer as one launch (one accumulator buffer: the MMA waits for the drain)
vs two 128-column launches (two buffers: drain overlaps the next tile's MMA, but the rows are read twice)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssf_slam_b200 import functional as F_, tc

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

g = torch.Generator(device="cuda").manual_seed(0)
for rows, K, N in [(262144, 256, 256), (524288, 128, 256), (131072, 256, 512), (524288, 256, 256)]:
    X = torch.randn(rows, K, device="cuda", generator=g)
    W = torch.randn(N, K, generator=torch.Generator().manual_seed(1)) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    img = tc.dense_image(W).cuda()
    halves = [(tc.dense_image(W[n0:n0 + 128]).cuda(), b[n0:n0 + 128].contiguous()) for n0 in range(0, N, 128)]
    one = t(lambda: F_.dense_tc(img, N, K, x1=X, bias=b, act=2))
    two = t(lambda: [F_.dense_tc(im, 128, K, x1=X, bias=bb, act=2) for im, bb in halves])
    print("rows=%d K=%d N=%d: one launch %.3f ms, %d launches of 128 columns %.3f ms" % (rows, K, N, one, len(halves), two))
