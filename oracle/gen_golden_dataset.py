"""Generates tests/golden/carla_subsample.npz by running the UNMODIFIED reference method
``CARLA3D.subsample_points`` (ASF/utils/datasets/carla.py:202-305) on a synthetic raw frame, for two flag settings.
Run in the build container only (needs /root/reference):  python oracle/gen_golden_dataset.py"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path[:0] = [ROOT, "/root/reference/scripts/ActiveSceneFlow"]
from utils.datasets import carla as ref_carla  # noqa: E402  (reference, unmodified)


def raw_frame(seed, n1, n2):
    rng = np.random.default_rng(seed)
    pos1 = rng.uniform(-40, 40, (n1, 3)).astype(np.float32)
    pos1[:, 2] = rng.uniform(-4.5, 3.0, n1)
    pos2 = rng.uniform(-40, 40, (n2, 3)).astype(np.float32)
    pos2[:, 2] = rng.uniform(-4.5, 3.0, n2)
    return dict(pos1=pos1, pos2=pos2, ego_flow=rng.normal(size=(n1, 3)).astype(np.float32),
                gt=rng.normal(size=(n1, 3)).astype(np.float32), s_fg_mask=(rng.random(n1) < 0.3).astype(np.float64),
                t_fg_mask=(rng.random(n2) < 0.3).astype(np.float64))


def run_reference(frame, nb_points, seed, **flags):
    me = types.SimpleNamespace(nb_points=nb_points, rm_ground=False, use_fg_inds=True, hybrid_sample=False, pre_segfrnt=True)
    me.__dict__.update(flags)
    cls = [c for c in vars(ref_carla).values() if isinstance(c, type) and hasattr(c, "subsample_points")][0]
    me.hybrid_sample_points = types.MethodType(cls.hybrid_sample_points, me)
    seq = [frame["pos1"].copy(), frame["pos2"].copy()]
    gt = [frame["ego_flow"].copy(), frame["gt"].copy()]
    mask = [frame["s_fg_mask"].copy(), frame["t_fg_mask"].copy()]
    np.random.seed(seed)
    s, g, m = cls.subsample_points(me, seq, gt, mask)
    return s, g, m


def main():
    out = {}
    frame = raw_frame(0, 3000, 2800)
    for tag, nb, flags in (("default", 512, {}), ("noseg_rmground", 1024, dict(pre_segfrnt=False, rm_ground=True)),
                           ("small_replace", 1024, dict(pre_segfrnt=True)), ("hybrid", 1024, dict(hybrid_sample=True))):
        s, g, m = run_reference(frame, nb, 1234, **flags)
        out.update({tag + "_pos1": s[0], tag + "_pos2": s[1], tag + "_ego": g[0], tag + "_gt": g[1], tag + "_m0": m[0], tag + "_m1": m[1]})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "carla_subsample.npz"), **{k: frame[k] for k in frame}, **out)
    print("wrote", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
