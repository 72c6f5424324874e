"""Writes tests/golden/gmm_mask.npz: inputs, the labels scikit-learn's own GaussianMixture produces from the specification's
initial parameters, and the specification's (oracle/gmm.py) results, after asserting that the two agree exactly.
Run in the build container (scikit-learn 1.9.0):  python -m oracle.gen_golden_gmm"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gmm  # noqa: E402
from ssf_slam_b200 import synth  # noqa: E402


def cases():
    """(name, points, flow): synthetic street frames driven with ground-truth flow + 2 cm noise (as the GT variants of the
    reference do), and one with network-like tiny flow."""
    out = []
    for seed, n in ((0, 2048), (3, 4096), (7, 2048)):
        it = synth.make_sequence(seed, 1, n)[0]
        rng = np.random.default_rng(100 + seed)
        out.append(("gt%d" % seed, it["pos1"], (it["gt"] + rng.normal(0, 0.02, it["gt"].shape)).astype(np.float32)))
    it = synth.make_sequence(11, 1, 2048)[0]
    out.append(("tiny", it["pos1"], (np.random.default_rng(5).normal(0, 0.03, it["gt"].shape)).astype(np.float32)))
    return out


def main():
    store = {}
    for name, p, f in cases():
        spec, sk_labels, gm = gmm.check_against_sklearn(p, f)
        assert np.array_equal(sk_labels, spec["labels"]), name
        assert gm.n_iter_ == spec["n_iter"] and abs(gm.lower_bound_ - spec["lower_bound"]) < 1e-9, name
        assert np.array_equal(gmm.reference_bg_index(sk_labels), spec["bg_index"]), name
        store.update({name + "_points": p, name + "_flow": f, name + "_sklearn_labels": sk_labels.astype(np.int8),
                      name + "_mask": spec["mask"], name + "_n_iter": np.int64(spec["n_iter"]),
                      name + "_lower_bound": np.float64(spec["lower_bound"]), name + "_sklearn_lower_bound": np.float64(gm.lower_bound_)})
        print(name, "n_iter", spec["n_iter"], "bg", len(spec["bg_index"]), "of", len(p), "min margin %.3g" % spec["margin"].min())
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gmm_mask.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
