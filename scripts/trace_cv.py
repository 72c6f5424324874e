"""Developer tool: phase timeline of the tensor-core cost volume (needs the library built with -DSSF_CV_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ssf_slam_b200 import functional as F_, synth, _native as nat
from ssf_slam_b200.model import prepare_weights
from ssf_slam_b200.weights import random_init_state_dict
B = 4
pool = synth.make_sequence(1000, B, 8192)
x1 = torch.from_numpy(np.stack([it["pos1"] for it in pool])).cuda()
x2 = torch.from_numpy(np.stack([it["pos2"] for it in pool])).cuda()
W = prepare_weights(random_init_state_dict(0), torch.device("cuda:0"))
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device="cuda")
idx = F_.knn_idx(16, x1, x2); idxw = F_.knn_idx(16, x2, x1)
Gab, Hab, H3 = r(B, 8192, 128) * 0.7, r(B, 8192, 128) * 0.7, r(B, 8192, 64) * 0.5
for _ in range(3):
    F_.cost_volume(Gab, Hab, W["flow0_r"], H3, x1, x2, idx, idxw, 64)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (3 * 8 * 32))()
L = nat.lib()
L.raw.ssf_cv_trace_read.argtypes = [ctypes.c_void_p]
assert L.raw.ssf_cv_trace_read(buf) == 0
t = np.array(buf[:]).reshape(3, 8, 32)
t0 = t[0, 0, 0]
names = ["start", "P done", "E1 wait", "E1 done", "bar", "ATT done", "E2 wait", "E2 done", "E3 wait", "E3 done", "E4 wait", "E4 done", "E5 wait", "E5 done"]
for it in range(2, 4):
    print("tile", it)
    for who, nm in ((0, "fwd"), (1, "wrp")):
        print("  %s: " % nm + "  ".join("%s=%d" % (n, t[who, it, i] - t0) for i, n in enumerate(names)))
    for who, nm in ((0, "fwd"), (1, "wrp")):
        print("  %s ATT: " % nm + "  ".join("%s=%d" % (n, t[who, it, 14 + i] - t0) for i, n in enumerate(["Q done", "bar", "N done", "S done+bar", "mix done", "bar"])))
    print("  mma: " + "  ".join("s%d%s %d-%d" % (st, "aw"[br], t[2, it, st * 4 + br * 2] - t0, t[2, it, st * 4 + br * 2 + 1] - t0) for st in range(5) for br in range(2)))
