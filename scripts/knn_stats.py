"""Developer tool (SSF_CV_TRACE=1 build): per-query visit / insert counts of the block kNN on realistic clouds."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ssf_slam_b200 import functional as F_, synth, _native as nat
B = 4
pool = synth.make_sequence(1000, B, 8192)
x1 = torch.from_numpy(np.stack([it["pos1"] for it in pool])).cuda()
x2 = torch.from_numpy(np.stack([it["pos2"] for it in pool])).cuda()
sub = F_.gather_rows(x1, F_.fps(x1, 2048))
L = nat.lib()
L.raw.ssf_knn_stat_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * 8)()
def stat(tag, fn):
    torch.cuda.synchronize(); L.raw.ssf_knn_stat_read(buf, 1)
    F_.knn_cache_clear(); fn(); torch.cuda.synchronize(); L.raw.ssf_knn_stat_read(buf, 1)
    q, v, ins, rc = buf[0], buf[1], buf[2], buf[3]
    print("%-28s queries %8d  visits/query %6.2f  inserts/query %6.2f (merge = 8)  sweep re-checks/query %6.2f  merges/query %5.2f  survivors/query %6.2f" % (
        tag, q, v / q, ins / q, rc / q, buf[4] / q, buf[5] / q))
stat("k16 8192x8192", lambda: F_.knn_idx(16, x1, x2))
stat("k7  8192x8192", lambda: F_.knn_idx(7, x1, x2))
stat("k16 8192x2048", lambda: F_.knn_idx(16, x1, sub))
stat("k16 2048x8192", lambda: F_.knn_idx(16, sub, x1))
stat("k16 2048x2048", lambda: F_.knn_idx(16, sub, sub))
