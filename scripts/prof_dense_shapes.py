"""Developer tool: plain-row dense_tc launches at the shapes of the B = 64 step, timed with CUDA events; the last launch of each
shape sits between cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ssf_slam_b200 import functional as F_, tc

shapes = [(262144, 256, 256), (524288, 128, 128), (524288, 64, 64), (1048576, 96, 64), (131072, 256, 512), (32768, 512, 256)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
g = torch.Generator(device="cuda").manual_seed(0)
for rows, K, N in shapes:
    X = torch.randn(rows, K, device="cuda", generator=g)
    W = torch.randn(N, K, generator=torch.Generator().manual_seed(1)) / K ** 0.5
    img = tc.dense_image(W).cuda()
    b = torch.randn(N, device="cuda", generator=g)
    for _ in range(3):
        F_.dense_tc(img, N, K, x1=X, bias=b, act=2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        F_.dense_tc(img, N, K, x1=X, bias=b, act=2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles = (rows + 127) // 128
    per_cta = -(-tiles // 148)
    print("rows=%d K=%d N=%d: %.3f ms, %.1f TFLOP/s algorithmic, %.0f cycles per tile per CTA, %.2f TB/s in+out" %
          (rows, K, N, ms, 2.0 * rows * K * N / ms / 1e9, ms * 1e-3 * 1.965e9 / per_cta / max(1, N // 256), rows * (K + N) * 4 / ms / 1e9))
    torch.cuda.profiler.start()
    F_.dense_tc(img, N, K, x1=X, bias=b, act=2)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
