"""oracle/gen_golden_point_twins.py -- TEST INFRASTRUCTURE.  Run in the BUILD container only (needs /root/reference).

Pins the point-operator oracle (oracle/point_ops.py, oracle/c/ssf_oracle.c) against the pure-PyTorch twins of the
pointnet2 extension that the reference ships in its own tree, executed UNMODIFIED:

* ``farthest_point_sample``  ASF/utils/utils.py:68-89   (imported; its ``torch.randint`` start index is forced to 0,
                                                         the extension's start -- the only thing patched, from outside)
* ``knn_point``              ASF/utils/utils.py:92-108  (imported)
* ``query_ball_point``       ASF/SetCover.py:39-63 with ``square_distance`` :17-37.  SetCover.py cannot be imported
  (``from HPR import *`` needs pyhull) and the function's last two statements (``dists = sqrdists[group_idx]``,
  ``mask2 = 1 - mask``) raise on any current torch, so the source text of both functions is ``exec``'d verbatim from the
  file and the function body is cut after ``group_idx[mask] = group_first[mask]`` -- the statements that define the
  indices and the counts run exactly as written.

The twins compute distances with other arithmetic than the extension (``sum((a-b)**2)`` over a repeated tensor, or the
``-2ab + a^2 + b^2`` expansion), so the clouds are TIE-FREE with margins: no two candidate distances of a query are closer
than the twins' rounding error, no pair distance sits within the error of a ball radius, and FPS maxima are unique by
a margin.  On such data every implementation of the mathematical definition must give the same indices; the oracle
(C and torch restatements) is asserted equal to the reference twins and the vectors are committed to
tests/golden/point_twins.npz, where the CPU tests re-check the oracle and the GPU tests check the CUDA operators.

Usage:  python -m oracle.gen_golden_point_twins
"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import point_ops  # noqa: E402
from oracle.ref_harness import REF_ASF, import_reference_tflow  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def reference_utils():
    import_reference_tflow()           # sets up sys.path (reference + shims)
    import utils.utils as ref_utils    # the reference's ASF/utils/utils.py, unmodified
    return ref_utils


def reference_ball_query():
    path = os.path.join(REF_ASF, "SetCover.py")
    tree = ast.parse(open(path).read())
    fns = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("square_distance", "query_ball_point")}
    qb = fns["query_ball_point"]
    cut = [i for i, st in enumerate(qb.body) if isinstance(st, ast.Assign) and getattr(st.targets[0], "id", None) == "dists"][0]
    qb.body = qb.body[:cut] + [ast.parse("return group_idx, cnt").body[0]]
    mod = ast.Module([fns["square_distance"], qb], [])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch}
    exec(compile(mod, path, "exec"), ns)
    return ns["query_ball_point"]


def _pair_d2(q, r):
    return ((q[:, :, None, :].astype(np.float64) - r[:, None, :, :].astype(np.float64)) ** 2).sum(-1)


def tie_free_queries(q, r, k, rel_gap):
    """Rows of q whose k+1 smallest distances to r are separated by more than rel_gap (relative)."""
    d2 = np.sort(_pair_d2(q, r), axis=-1)[:, :, :k + 1]
    gap = (d2[:, :, 1:] - d2[:, :, :-1]) / np.maximum(d2[:, :, 1:], 1e-12)
    ok = (gap > rel_gap).all(-1)
    return ok.all(0)   # the same query rows in every cloud of the batch


def main():
    ru = reference_utils()
    rng = np.random.default_rng(20250)
    out = {}

    # ---- FPS: 2 clouds, 4096 -> 512; the cloud seed is searched until every argmax is unique by a margin of 5e-6 relative
    # (fp32 rounding of a squared distance is ~3e-7), checked in fp64
    real_randint = torch.randint
    for fps_seed in range(100):
        r2 = np.random.default_rng(1000 + fps_seed)
        xyz = (r2.uniform(-1, 1, (2, 4096, 3)) * np.array([30.0, 20.0, 2.0])).astype(np.float32)
        torch.randint = lambda lo, hi, size, **kw: torch.zeros(size, dtype=kw.get("dtype", torch.long))
        try:
            ref_fps = ru.farthest_point_sample(torch.from_numpy(xyz), 512).numpy()
        finally:
            torch.randint = real_randint
        mind = np.full(xyz.shape[:2], 1e10)
        clean = True
        for j in range(511):
            c = xyz[np.arange(2), ref_fps[:, j]].astype(np.float64)
            mind = np.minimum(mind, ((xyz.astype(np.float64) - c[:, None]) ** 2).sum(-1))
            top2 = np.sort(mind, axis=1)[:, -2:]
            if not ((top2[:, 1] - top2[:, 0]) > 5e-6 * top2[:, 1]).all():
                clean = False
                break
            assert (np.argmax(mind, axis=1) == ref_fps[:, j + 1]).all()
        if clean:
            break
    assert clean, "no tie-free FPS cloud found"
    out["fps_seed"] = 1000 + fps_seed
    for use_c in (True, False):
        point_ops.USE_C = use_c
        got = point_ops.furthest_point_sample(torch.from_numpy(xyz), 512).numpy()
        assert np.array_equal(got, ref_fps.astype(np.int32)), "oracle FPS (C=%s) deviates from the reference twin" % use_c
    out.update(fps_xyz=xyz, fps_idx=ref_fps.astype(np.int32))
    print("FPS 4096->512 x2: reference farthest_point_sample == oracle (C and torch)")

    # ---- kNN: reference knn_point(k, dataset, queries); queries filtered to tie-free rows
    ref_pts = (rng.uniform(-1, 1, (2, 3000, 3)) * np.array([30.0, 20.0, 2.0])).astype(np.float32)
    q = (rng.uniform(-1, 1, (2, 1500, 3)) * np.array([30.0, 20.0, 2.0])).astype(np.float32)
    keep = tie_free_queries(q, ref_pts, 16, 1e-4)
    q = np.ascontiguousarray(q[:, keep][:, :1024])
    assert q.shape[1] == 1024
    for k in (3, 8, 16):
        val, idx = ru.knn_point(k, torch.from_numpy(ref_pts), torch.from_numpy(q))
        for use_c in (True, False):
            point_ops.USE_C = use_c
            d, i = point_ops.knn(k, torch.from_numpy(q), torch.from_numpy(ref_pts))
            assert np.array_equal(i.numpy(), idx.numpy().astype(np.int32)), "oracle kNN k=%d (C=%s) deviates" % (k, use_c)
            assert np.allclose(d.numpy(), val.numpy(), rtol=1e-5, atol=1e-6)
        out["knn%d_idx" % k] = idx.numpy().astype(np.int32)
        out["knn%d_dist" % k] = val.numpy()
    out.update(knn_ref=ref_pts, knn_query=q)
    print("kNN k=3/8/16, 1024 x 3000 x2: reference knn_point == oracle (C and torch)")

    # ---- ball query: SetCover.query_ball_point (statements up to the index/count definition, verbatim)
    qb = reference_ball_query()
    cloud = rng.uniform(-4, 4, (2, 2048, 3)).astype(np.float32)
    cent = rng.uniform(-4, 4, (2, 600, 3)).astype(np.float32)
    d2 = _pair_d2(cent, cloud)
    radii = (0.5, 1.0, 2.0, 4.0)
    ok = np.ones(cent.shape[1], bool)
    for r in radii:
        ok &= (np.abs(d2 - r * r) > 2e-4).all(-1).all(0)   # |x|^2 <= 48: the expansion's error is ~1e-5
    cent = np.ascontiguousarray(cent[:, ok][:, :256])
    assert cent.shape[1] == 256
    for r in radii:
        gi, cnt = qb(r, 16, torch.from_numpy(cloud), torch.from_numpy(cent))
        gi, cnt = gi.numpy(), cnt.numpy()
        empty = cnt == 0
        # a query with no hit: the reference leaves the sentinel N in every slot (an out-of-range index); the extension's
        # convention (and the oracle's) is zeros.  Only rows with hits are compared; empty rows are recorded.
        oi, oc = point_ops.ball_query(r, 16, torch.from_numpy(cloud), torch.from_numpy(cent))
        ci, cc = point_ops.c_ball_query(r, 16, cloud, cent)
        assert np.array_equal(oc.numpy(), cnt) and np.array_equal(cc, cnt), "ball query counts r=%g" % r
        assert np.array_equal(oi.numpy()[~empty], gi[~empty]) and np.array_equal(ci[~empty], gi[~empty]), "ball query idx r=%g" % r
        assert (gi[empty] == cloud.shape[1]).all() and (oi.numpy()[empty] == 0).all()
        gi = gi.copy()
        gi[empty] = 0
        out["ball_r%g_idx" % r] = gi.astype(np.int32)
        out["ball_r%g_cnt" % r] = cnt.astype(np.int32)
        print("ball query r=%g: reference SetCover.query_ball_point == oracle (hits/query: mean %.1f, %d empty)" %
              (r, cnt.mean(), int(empty.sum())))
    out.update(ball_xyz=cloud, ball_new_xyz=cent)
    point_ops.USE_C = True
    np.savez_compressed(os.path.join(OUT, "point_twins.npz"), **out)
    print("wrote tests/golden/point_twins.npz")


if __name__ == "__main__":
    main()
