"""TEST INFRASTRUCTURE (oracle): plane-feature extraction of src/frameFeature.cpp:45-127.

Two independent restatements that must agree: ``plane_features_c`` (oracle/c/ssf_oracle.c, line-by-line C) and
``plane_features_py`` (plain Python/NumPy fp32 loops, small inputs only).  PINNED against the reference's own node: its
src/frameFeature.cpp compiles unmodified, from where it lies, against stand-in ROS / PCL headers (oracle/ref_build/ ->
oracle/_ref/libframe_feature_ref.so); oracle/gen_golden_plane_features.py asserts both restatements equal what that node
publishes, bit for bit, and commits the vectors (tests/golden/plane_features_ref.npz)."""
import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = {16: (0.05, 3, 0, 0), 64: (0.005, 25, 5, 5)}


def plane_features_c(points, n_rows=16, plane_min=None, plane_span=None, row_start=None, row_end=None):
    d = PARAMS[n_rows]
    plane_min, plane_span = (d[0] if plane_min is None else plane_min), (d[1] if plane_span is None else plane_span)
    row_start, row_end = (d[2] if row_start is None else row_start), (d[3] if row_end is None else row_end)
    lib = ctypes.CDLL(os.path.join(_HERE, "c", "libssf_oracle.so"))
    fn = lib.ssf_oracle_plane_features
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    P = np.ascontiguousarray(points, np.float32)
    out = np.zeros((P.shape[0], 4), np.float32)
    n = fn(P.ctypes.data, P.shape[0], n_rows, row_start, row_end, plane_min, plane_span, out.ctypes.data)
    return out[:n].copy()


def plane_features_py(points, n_rows=16, plane_min=None, plane_span=None, row_start=None, row_end=None):
    d = PARAMS[n_rows]
    plane_min = np.float32(d[0] if plane_min is None else plane_min)
    plane_span = d[1] if plane_span is None else plane_span
    row_start, row_end = (d[2] if row_start is None else row_start), (d[3] if row_end is None else row_end)
    f = np.float32
    P = np.asarray(points, np.float32)
    rows = [[] for _ in range(n_rows)]
    for i in range(P.shape[0]):
        x, y, z = P[i]
        with np.errstate(divide="ignore", invalid="ignore"):
            a = f(math.atan(float(f(z / np.sqrt(f(x * x + y * y))))))   # atanf: double atan rounded to float
        angle = f(float(f(a * f(180))) / math.pi)
        if np.isnan(angle):
            continue
        sid = -1
        if n_rows == 16:
            sid = int(float(f(f(angle + f(15)) / f(2))) + 0.5)
        else:
            sid = int(float(f(f(2) - angle)) * 3.0 + 0.5) if float(angle) >= -8.83 else n_rows // 2 + int((-8.83 - float(angle)) * 2.0 + 0.5)
        if -1 < sid < n_rows:
            rows[sid].append(i)
    out = []
    for r in range(row_start, n_rows - row_end):
        s = rows[r]
        value = np.zeros(len(s), np.float32)
        for j in range(5, len(s) - 5):
            dd = []
            for c in range(3):
                v = lambda k: P[s[j + k], c]
                acc = f(v(-5) + v(-4))
                for k in (-3, -2, -1):
                    acc = f(acc + v(k))
                acc = f(acc - f(f(10) * v(0)))
                for k in (1, 2, 3, 4, 5):
                    acc = f(acc + v(k))
                dd.append(acc)
            value[j] = f(f(f(dd[0] * dd[0]) + f(dd[1] * dd[1])) + f(dd[2] * dd[2]))
        jstart = 0
        for j in range(len(s)):
            if j >= jstart and value[j] < plane_min:
                out.append([P[s[j], 0], P[s[j], 1], P[s[j], 2], f(j + r / 100.0)])
                jstart = j + plane_span
    return np.asarray(out, np.float32).reshape(-1, 4)
