"""One launch of every stand-alone point operator (and of the hot path's HBM-bound helpers) at the shapes bench.py's `point_ops`
leg times, between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv ...
`--order-out` writes the op label of every profiled launch, in launch order, for scripts/ncu_traffic.py."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ssf_slam_b200 import _native as nat, functional as F_, pointnet2_utils as pu

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--npoints", type=int, default=8192)
ap.add_argument("--order-out", default=None)
args = ap.parse_args()
B, N, M = args.batch, args.npoints, 2048
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
xyz = torch.randn(B, N, 3, device=dev, generator=g) * torch.tensor([30.0, 20.0, 2.0], device=dev)
fps_idx = pu.furthest_point_sample(xyz, M)
new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
_, idx16 = pu.knn(16, xyz, xyz)
feat96 = torch.randn(B, 96, N, device=dev, generator=g)
w3 = torch.rand(B, N, 3, device=dev, generator=g)
idx3 = idx16[:, :, :3].contiguous()
val64 = torch.randn(B, N, 64, device=dev, generator=g)          # point-major features for the model-side helpers
_, idx_up = pu.knn(16, xyz, new_xyz)
sparse_val = torch.randn(B, M, 64, device=dev, generator=g)
logit = torch.randn(B, N * 16, device=dev, generator=g)
rows = torch.randn(B, N * 16, 64, device=dev, generator=g)
key = idx16.view(B, N * 16).contiguous()

ops = [
    ("grouping_operation[C=96,M=%d,S=16]" % N, lambda: pu.grouping_operation(feat96, idx16)),
    ("three_interpolate[C=96,n=%d]" % N, lambda: pu.three_interpolate(feat96, idx3, w3)),
    ("gather_operation[C=96,M=%d]" % M, lambda: pu.gather_operation(feat96, fps_idx)),
    ("furthest_point_sample[N=%d,n=%d]" % (N, M), lambda: pu.furthest_point_sample(xyz, M)),
    ("knn[k=16,%dx%d]" % (N, N), lambda: pu.knn(16, xyz, xyz)),
    ("ball_query[r=1.0,ns=16,%dx%d]" % (M, N), lambda: pu.ball_query(1.0, 16, xyz, new_xyz)),
    ("interpolate[k=7,C=64,%d<-%d]" % (N, M), lambda: F_.interpolate(xyz, new_xyz, sparse_val, idx_up, mode=0, clampv=100.0, k=7)),
    ("build_csr[L=%d]" % (N * 16), lambda: F_.build_csr(key, N)),
]
csr = F_.build_csr(key, N)
ops.append(("segment_softmax_sum[L=%d,C=64]" % (N * 16), lambda: F_.segment_softmax_sum(logit, rows, csr, N)))

for _, fn in ops:       # warm-up (attributes, allocator)
    fn()
torch.cuda.synchronize()
order = []
torch.cuda.profiler.start()
for name, fn in ops:
    l0 = nat.launch_count()
    fn()
    torch.cuda.synchronize()
    order += [name] * (nat.launch_count() - l0)
torch.cuda.profiler.stop()
if args.order_out:
    json.dump({"batch": B, "order": order}, open(args.order_out, "w"))
print("profiled launches:", len(order))
