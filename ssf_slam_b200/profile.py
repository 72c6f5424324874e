"""Per-kernel CUDA-event timing of our launches (bench.py's kernel shares / roofline leg).  ``enable()`` wraps the
functions of ``ssf_slam_b200.functional`` so that every launch is bracketed by two events on the launching stream."""
import torch

from . import functional as F_

_NAMES = ("linear", "dense_tc", "attention_mix", "softmax_pool", "gather_rows", "transpose", "fps", "knn_idx", "interpolate", "group_mlp_max", "cost_volume",
          "build_csr", "segment_softmax_sum", "frontend", "gmm_mask")
_orig = {}
_records = []


def _tag(name, args, kwargs):
    try:
        if name == "cost_volume":
            return "cost_volume[N1=%d,m=%d]" % (args[4].shape[1], args[8])
        if name == "group_mlp_max":
            return "group_mlp_max[Nq=%d,S=%d,C1=%d]" % (args[1].shape[1], args[1].shape[2], args[0].shape[2])
        if name == "knn_idx":
            return "knn[k=%d,Nq=%d,Nr=%d]" % (args[0], args[1].shape[1], args[2].shape[1])
        if name == "dense_tc":
            rows = (kwargs["x1"].numel() // kwargs["x1"].shape[-1]) if kwargs.get("x1") is not None else kwargs["idx"].numel()
            mode = "g" if kwargs.get("G") is not None else "r"
            return "dense_tc[%s,rows=%d,K=%d,N=%d,epi=%d]" % (mode, rows, args[2], args[1], kwargs.get("epi", 0))
        if name == "fps":
            return "fps[N=%d,n=%d]" % (args[0].shape[1], args[1])
    except Exception:
        pass
    return name


def enable():
    if _orig:
        return
    for n in _NAMES:
        fn = getattr(F_, n)
        _orig[n] = fn

        def wrapped(*a, _fn=fn, _n=n, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = _fn(*a, **k)
            e1.record()
            _records.append((_tag(_n, a, k), e0, e1))
            return out

        setattr(F_, n, wrapped)


def disable():
    for n, fn in _orig.items():
        setattr(F_, n, fn)
    _orig.clear()
    _records.clear()


def summary():
    """{tag: {ms, calls, share}} over everything recorded since enable(); call after torch.cuda.synchronize()."""
    out = {}
    for tag, e0, e1 in _records:
        d = out.setdefault(tag, {"ms": 0.0, "calls": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["calls"] += 1
    total = sum(d["ms"] for d in out.values()) or 1.0
    for d in out.values():
        d["share"] = d["ms"] / total
    return out
