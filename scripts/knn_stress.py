"""Developer check: the block search on seeded random shapes (lattice clouds = massive distance ties included), results saved
for a bit-by-bit comparison between builds / switches (SSF_KNN_SB=0: register-resident bounds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ssf_slam_b200 import functional as F_, synth
rng = np.random.default_rng(321)
out = {}
for t in range(48):
    Nr = int(rng.integers(2049, 16385)); Nq = int(rng.integers(100, 5000)); B = int(rng.integers(1, 3)); k = int(rng.choice([1, 3, 7, 16, 32]))
    kind = t % 3
    if kind == 0:
        ref = rng.standard_normal((B, Nr, 3)).astype(np.float32) * np.array([40, 10, 2], np.float32)
        q = rng.standard_normal((B, Nq, 3)).astype(np.float32) * np.array([45, 12, 3], np.float32)
    elif kind == 1:
        ref = np.round(rng.standard_normal((B, Nr, 3)) * 4).astype(np.float32)
        q = np.round(rng.standard_normal((B, Nq, 3)) * 4).astype(np.float32)
    else:
        pool = synth.make_sequence(700 + t, B, 8192)
        ref = np.stack([it["pos2"][:min(Nr, 8192)] for it in pool]).astype(np.float32); q = np.stack([it["pos1"][:Nq] for it in pool]).astype(np.float32)
    F_.knn_cache_clear()
    out["t%d" % t] = F_.knn_idx(k, torch.from_numpy(np.ascontiguousarray(q)).cuda(), torch.from_numpy(np.ascontiguousarray(ref)).cuda()).cpu().numpy()
np.savez(sys.argv[1], **out)
print("done", len(out))
