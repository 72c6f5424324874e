"""GPU bring-up helper: structured inputs through ssf_tc_gemm_test to reveal operand layout mistakes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssf_slam_b200 import _native as nat, tc
nat.require_device()
torch.set_printoptions(linewidth=200, precision=4, sci_mode=False)

def run(X, W, mode, passes):
    K, N = X.shape[1], W.shape[0]
    hi, lo = tc.weight_image(W)
    Y = torch.full((128, N), float("nan"), device="cuda")
    Xd, hid, lod = X.cuda().contiguous(), hi.cuda(), lo.cuda()
    nat.check(nat.dev_lib().ssf_tc_gemm_test(nat.ptr(Xd), nat.ptr(hid), nat.ptr(lod), K, N, mode, passes, nat.ptr(Y), nat.stream()), nat.dev_lib())
    torch.cuda.synchronize()
    return Y.cpu()

for mode in (0, 1):
    K, N = 16, 16
    X = torch.zeros(128, K); W = torch.zeros(N, K)
    X[:, 0] = torch.arange(128).float() + 1; W[:, 0] = torch.arange(N).float() + 1
    Y = run(X, W, mode, 1)
    print("mode", mode, "k=0 only: Y[0:3,:8]", Y[0:3, :8].tolist(), "Y[100,:4]", Y[100, :4].tolist(), "nan", int(Y.isnan().sum()))
    for kk in (1, 3, 4, 5, 8, 13):
        X = torch.zeros(128, K); W = torch.zeros(N, K)
        X[:, kk] = 1.0; W[:, kk] = torch.arange(N).float() + 1
        Y = run(X, W, mode, 1)
        print("  k=%d: Y[0,:8]" % kk, Y[0, :8].tolist(), "Y[9,:4]", Y[9, :4].tolist())
    X = torch.ones(128, K); W = torch.ones(N, K)
    print("  ones: ", run(X, W, mode, 1)[0, :4].tolist(), "expect", K)
    g = torch.Generator().manual_seed(0)
    X = torch.randn(128, 64, generator=g); W = torch.randn(64, 64, generator=g)
    Y = run(X, W, mode, 3); ref = X @ W.t()
    print("  rand: Y[0,:6]", Y[0, :6].tolist()); print("        ref   ", ref[0, :6].tolist())
    print("  |Y| max %.4g mean %.4g ; |ref| max %.4g" % (Y.abs().max(), Y.abs().mean(), ref.abs().max()))

import numpy as np
os.makedirs("gpurun_out", exist_ok=True)
g = torch.Generator().manual_seed(0)
X = torch.randn(128, 64, generator=g); W = torch.randn(64, 64, generator=g)
out = {"X": X.numpy(), "W": W.numpy()}
for mode in (0, 1):
    for passes in (1, 3):
        out["Y_m%d_p%d" % (mode, passes)] = run(X, W, mode, passes).numpy()
np.savez("gpurun_out/tc_debug.npz", **out)
