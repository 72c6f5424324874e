"""Developer tool (SSF_CV_TRACE=1 build): where a producer warp and an epilogue warp of dense_tc spend their cycles
(CTA 0, clock64 deltas accumulated per phase), for plain-row layers `rows x K x N` given on the command line."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssf_slam_b200 import functional as F_, tc, _native as nat

L = nat.lib()
L.raw.ssf_dense_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_longlong * 48)()
shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:] if not s.startswith("g")] or [(524288, 64, 64), (262144, 256, 256)]
grouped = [a for a in sys.argv[1:] if a.startswith("g")]   # e.g. g8x8192x2048x64: B clouds, Nq query points, Nsrc source rows, K = N
for rows, K, N in shapes:
    X = torch.randn(rows, K, device="cuda")
    img = tc.dense_image(torch.randn(N, K) / K ** 0.5).cuda()
    b = torch.randn(N, device="cuda")
    for _ in range(2):
        F_.dense_tc(img, N, K, x1=X, bias=b, act=2)
    torch.cuda.synchronize()
    L.raw.ssf_dense_trace_read(buf, 1)
    F_.dense_tc(img, N, K, x1=X, bias=b, act=2)
    torch.cuda.synchronize()
    L.raw.ssf_dense_trace_read(buf, 1)
    t = list(buf)
    tiles = -(-((rows + 127) // 128) // 148)     # tiles of CTA 0
    chunks = tiles * (K // 32)
    print("rows=%d K=%d N=%d: %d tiles, %d chunks in CTA 0" % (rows, K, N, tiles, chunks))
    names = ["loop rest", "issue cp.async", "wait+read chunk", "first-layer math", "wait A stage", "split + st issue", "wait::st", "", "", "", "", "", "", "", "", "arrive"]
    print("  producer (cycles per chunk): " + "  ".join("%s=%.0f" % (names[i], t[i] / chunks) for i in (0, 1, 2, 3, 4, 5, 6, 15)))
    print("    inside the prefetch issue (tensor-map path): address arithmetic=%.0f  expect_tx=%.0f  cp.async.bulk.tensor issue=%.0f" % tuple(t[i] / chunks for i in (7, 8, 9)))
    print("  epilogue (cycles per tile):  setup=%.0f  wait accumulator=%.0f  column loop=%.0f  arrive=%.0f" %
          tuple(t[16 + i] / tiles / max(1, N // 256) for i in (0, 1, 2, 15)))
    print("    column loop per tile: tail=%.0f  tcgen05.ld+wait=%.0f  math=%.0f  staging stores=%.0f  row-segment stores=%.0f" %
          tuple(t[16 + i] / tiles / max(1, N // 256) for i in (3, 4, 5, 6, 7)))

for spec in grouped:
    B, Nq, Nsrc, K = (int(v) for v in spec[1:].split("x"))
    N = K
    G = torch.randn(B, Nsrc, K, device="cuda")
    ps, pq = torch.randn(B, Nsrc, 3, device="cuda"), torch.randn(B, Nq, 3, device="cuda")
    idx = F_.knn_idx(16, pq, ps)
    img = tc.dense_image(torch.randn(N, K) / K ** 0.5).cuda()
    b1, Wd1, b = torch.randn(K, device="cuda"), torch.randn(3, K, device="cuda"), torch.randn(N, device="cuda")
    run = lambda: F_.dense_tc(img, N, K, G=G, b1=b1, Wd1=Wd1, act1=1, idx=idx, pos_src=ps, pos_q=pq, bias=b, act=1, epi=F_.EPI_MAX)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    L.raw.ssf_dense_trace_read(buf, 1)
    run()
    torch.cuda.synchronize()
    L.raw.ssf_dense_trace_read(buf, 1)
    t = list(buf)
    tiles = -(-(B * Nq * 16 // 128) // 148)
    chunks = tiles * (K // 32)
    print("grouped B=%d Nq=%d Nsrc=%d K=N=%d: %d tiles, %d chunks in CTA 0" % (B, Nq, Nsrc, K, tiles, chunks))
    print("  producer (cycles per chunk): " + "  ".join("%s=%.0f" % (names[i], t[i] / chunks) for i in (0, 1, 2, 3, 4, 5, 6, 15)))
    print("  epilogue (cycles per tile, summed over the epilogue warpgroups):  setup=%.0f  wait accumulator=%.0f  column loop=%.0f  arrive=%.0f" %
          tuple(t[16 + i] / tiles for i in (0, 1, 2, 15)))
