"""Developer check: furthest point sampling on seeded random shapes (clusters, LiDAR sweeps, lattices = massive ties), results
saved for a bit-by-bit comparison between switches (SSF_FPS_PRUNE=0: the plain kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ssf_slam_b200 import functional as F_, synth
rng = np.random.default_rng(123)
out = {}
for t in range(60):
    N = int(rng.integers(2049, 8193)); B = int(rng.integers(1, 4)); npnt = int(rng.integers(1, N + 1))
    kind = t % 4
    if kind == 0:
        x = rng.standard_normal((B, N, 3)).astype(np.float32) * np.array([40, 10, 2], np.float32)
    elif kind == 1:
        c = rng.uniform(-60, 60, (B, 20, 1, 3)); x = (c + rng.standard_normal((B, 20, (N + 19) // 20, 3)) * 0.3).reshape(B, -1, 3)[:, :N].astype(np.float32)
    elif kind == 2:
        x = np.stack([synth.make_sequence(500 + t, 1, 8192)[0]["pos1"][:N] for _ in range(B)]).astype(np.float32)
    else:
        x = np.round(rng.standard_normal((B, N, 3)) * 3).astype(np.float32)      # lattice: massive distance ties
    out["t%d" % t] = F_.fps(torch.from_numpy(np.ascontiguousarray(x)).cuda(), npnt).cpu().numpy()
np.savez(sys.argv[1], **out)
print("done", len(out))
