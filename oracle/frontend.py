"""oracle/frontend.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the per-frame front end that follows the scene-flow network:

* ``solve_rt_svd``      -- the reference's ``slove_RT_by_SVD`` (scripts/PointCloudOdometry.py:15-33;
                            copies at PointCloudOdometry_noSeg.py:19-37, ASF/main_sju_occ_ros.py:455-473)
                            with the *intended* ``Vt.T @ U.T`` in the reflection branch (the shipped
                            ``&`` raises TypeError -- SURVEY.md Appendix C-13).
* ``rotation_to_quaternion`` -- what ``pyquaternion.Quaternion(matrix=R)`` yields (pyquaternion is an
                            unpinned, uninstalled dependency: requirement.txt:4); message order
                            ``[x,y,z,w]`` as in scripts/PointCloudOdometry.py:97-98.
* ``gmm_background``    -- the reference noSeg masker: 2-component sklearn GMM on [flow|xyz], majority
                            label = background (scripts/PointCloudOdometry_noSeg.py:97-103,
                            ASF/main_sju_occ_ros.py:257-263).  Not bit-reproducible (global NumPy RNG).
* ``gt_background``     -- ``argwhere(s_fg_mask == 0)`` (scripts/PointCloudOdometry.py:91).
* ``masker_spec`` / ``kabsch_spec`` -- the deterministic residual-vs-rigid-flow masker with
                            per-instance voting and weighted Kabsch defined in SURVEY.md Appendix D and
                            DESIGN.md (anchors: ASF/calc_coarse_flow.py:587-590 residual test,
                            scripts/PointCloudOdometry.py:93-96 orientation).  This is the bit-level
                            specification the CUDA kernels must reproduce: every reduction order,
                            rounding and comparison below is mirrored in csrc/frontend.cu.

PARITY UNPINNED for masker_spec (the reference has no such masker); solve_rt_svd is checked against
the reference's own function text executed verbatim in tests/test_oracle.py when /root/reference is
mounted.
"""
from collections import Counter

import numpy as np

# ------------------------------------------------------------------ reference restatements


def solve_rt_svd(src, dst):
    """(R [3,3], t [3,1]) with dst ~= R @ src + t; un-weighted Kabsch as the reference computes it."""
    src_mean = src.mean(axis=0, keepdims=True)
    dst_mean = dst.mean(axis=0, keepdims=True)
    a = src - src_mean
    b = dst - dst_mean
    H = a.T @ b
    U, _, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt = Vt.copy()
        Vt[2, :] *= -1
        R = Vt.T @ U.T
    t = -R @ src_mean.T + dst_mean.T
    return R, t


def rotation_to_quaternion(R):
    """[w, x, y, z] by the trace method pyquaternion uses (largest-diagonal branch selection)."""
    m = np.asarray(R, np.float64).T  # pyquaternion works on the transpose
    if m[2, 2] < 0:
        if m[0, 0] > m[1, 1]:
            t = 1 + m[0, 0] - m[1, 1] - m[2, 2]
            q = [m[1, 2] - m[2, 1], t, m[0, 1] + m[1, 0], m[2, 0] + m[0, 2]]
        else:
            t = 1 - m[0, 0] + m[1, 1] - m[2, 2]
            q = [m[2, 0] - m[0, 2], m[0, 1] + m[1, 0], t, m[1, 2] + m[2, 1]]
    else:
        if m[0, 0] < -m[1, 1]:
            t = 1 - m[0, 0] - m[1, 1] + m[2, 2]
            q = [m[0, 1] - m[1, 0], m[2, 0] + m[0, 2], m[1, 2] + m[2, 1], t]
        else:
            t = 1 + m[0, 0] + m[1, 1] + m[2, 2]
            q = [t, m[1, 2] - m[2, 1], m[2, 0] - m[0, 2], m[0, 1] - m[1, 0]]
    return np.array(q, np.float64) * (0.5 / np.sqrt(t))


def odom_message(R, t):
    """Float64MultiArray payload [tx,ty,tz,qx,qy,qz,qw] (scripts/PointCloudOdometry.py:97-103)."""
    w, x, y, z = rotation_to_quaternion(R)
    return np.hstack((np.asarray(t, np.float64).flatten(), [x, y, z, w]))


def gt_background(s_fg_mask):
    return np.argwhere(np.asarray(s_fg_mask) == 0).flatten()


def gmm_background(points, flow, random_state=None):
    from sklearn.mixture import GaussianMixture
    x = np.concatenate((flow, points), axis=1)
    labels = GaussianMixture(n_components=2, random_state=random_state).fit_predict(x)
    bg = Counter(labels).most_common(1)[0][0]
    return np.argwhere(labels == bg).flatten()


def reference_pose(points, flow, bg_index):
    """target = points+flow, source = points, slove_RT_by_SVD(target, source) (PointCloudOdometry.py:93-96)."""
    return solve_rt_svd(points[bg_index] + flow[bg_index], points[bg_index])


# ------------------------------------------------------------------ deterministic masker spec

FRONT_THREADS = 256  # one CTA of 256 threads per cloud; the reduction order below mirrors it
JACOBI_SWEEPS = 12


def _block_sum(v):
    """Order-exact mirror of the CUDA block reduction: v is fp64 [n, ...]; thread t adds elements
    t, t+256, ... sequentially; lanes combine with an xor butterfly (16,8,4,2,1); thread 0 then adds
    the 8 warp sums in warp order."""
    n = v.shape[0]
    T = FRONT_THREADS
    pad = (-n) % T
    if pad:
        v = np.concatenate([v, np.zeros((pad,) + v.shape[1:], v.dtype)], 0)
    v = v.reshape(-1, T, *v.shape[1:])
    acc = np.zeros(v.shape[1:], np.float64)
    for r in range(v.shape[0]):
        acc = acc + v[r]
    acc = acc.reshape(T // 32, 32, *acc.shape[1:])
    for off in (16, 8, 4, 2, 1):
        acc = acc + acc[:, np.arange(32) ^ off]
    total = acc[0, 0].copy()
    for w in range(1, T // 32):
        total = total + acc[w, 0]
    return total


def kabsch_sums(a32, b32, w):
    """16 fp64 sums: W, Sa[3], Sb[3], Sab[9] (row-major a x b).  a32, b32 fp32 [n,3]; w in {0,1}."""
    a = a32.astype(np.float64)
    b = b32.astype(np.float64)
    w = np.asarray(w, np.float64)
    cols = [w, w * a[:, 0], w * a[:, 1], w * a[:, 2], w * b[:, 0], w * b[:, 1], w * b[:, 2]]
    for i in range(3):
        for j in range(3):
            cols.append((w * a[:, i]) * b[:, j])
    return _block_sum(np.stack(cols, 1))


def _jacobi_eig4(A):
    """Cyclic Jacobi on a symmetric 4x4 (fp64), fixed JACOBI_SWEEPS sweeps, fixed pair order.
    Returns (eigenvalues diag, eigenvectors as columns).  Mirrors csrc/frontend.cu op for op."""
    A = A.copy()
    V = np.eye(4)
    old = np.seterr(over="ignore")  # theta*theta may overflow to inf once converged: t -> 0, as on the GPU
    try:
        _jacobi_sweeps(A, V)
    finally:
        np.seterr(**old)
    return np.diag(A).copy(), V


def _jacobi_sweeps(A, V):
    for _ in range(JACOBI_SWEEPS):
        for p in range(3):
            for q in range(p + 1, 4):
                apq = A[p, q]
                if apq == 0.0:
                    continue
                theta = (A[q, q] - A[p, p]) / (2.0 * apq)
                t = 1.0 / (abs(theta) + np.sqrt(theta * theta + 1.0))
                if theta < 0.0:
                    t = -t
                c = 1.0 / np.sqrt(t * t + 1.0)
                s = t * c
                for k in range(4):  # columns p, q of A
                    akp, akq = A[k, p], A[k, q]
                    A[k, p] = c * akp - s * akq
                    A[k, q] = s * akp + c * akq
                for k in range(4):  # rows p, q of A
                    apk, aqk = A[p, k], A[q, k]
                    A[p, k] = c * apk - s * aqk
                    A[q, k] = s * apk + c * aqk
                for k in range(4):
                    vkp, vkq = V[k, p], V[k, q]
                    V[k, p] = c * vkp - s * vkq
                    V[k, q] = s * vkp + c * vkq


def pose_from_sums(s):
    """Horn's closed form from the 16 sums: returns (quat [w,x,y,z] with w>=0, R fp64 [3,3], t fp64 [3]).
    b ~= R a + t.  Degenerate input (W < 3) -> identity."""
    W = s[0]
    if not W >= 3.0:
        return np.array([1.0, 0, 0, 0]), np.eye(3), np.zeros(3)
    ma = s[1:4] / W
    mb = s[4:7] / W
    S = s[7:16].reshape(3, 3) - W * np.outer(ma, mb)
    Sxx, Sxy, Sxz = S[0]
    Syx, Syy, Syz = S[1]
    Szx, Szy, Szz = S[2]
    Nm = np.array([
        [(Sxx + Syy) + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx],
        [Syz - Szy, (Sxx - Syy) - Szz, Sxy + Syx, Szx + Sxz],
        [Szx - Sxz, Sxy + Syx, (Syy - Sxx) - Szz, Syz + Szy],
        [Sxy - Syx, Szx + Sxz, Syz + Szy, (Szz - Sxx) - Syy]])
    ev, V = _jacobi_eig4(Nm)
    k = 0
    for i in range(1, 4):  # first maximum
        if ev[i] > ev[k]:
            k = i
    q = V[:, k].copy()
    n = np.sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3])
    q = q / n
    if q[0] < 0.0:
        q = -q
    w, x, y, z = q
    R = np.array([
        [1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - w * z), 2.0 * (x * z + w * y)],
        [2.0 * (x * y + w * z), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - w * x)],
        [2.0 * (x * z - w * y), 2.0 * (y * z + w * x), 1.0 - 2.0 * (x * x + y * y)]])
    t = np.array([mb[i] - ((R[i, 0] * ma[0] + R[i, 1] * ma[1]) + R[i, 2] * ma[2]) for i in range(3)])
    return q, R, t


def kabsch_spec(points, flow, weight):
    """Weighted Kabsch of (points+flow -> points); weight in {0,1}.  Returns (quat wxyz, R, t)."""
    p = np.asarray(points, np.float32)
    f = np.asarray(flow, np.float32)
    return pose_from_sums(kabsch_sums(p + f, p, weight))


def residual_sq(points, flow, R, t):
    """fp32 squared residual ||R(p+f)+t - p||^2 with every op rounded (no FMA)."""
    p = np.asarray(points, np.float32)
    q = p + np.asarray(flow, np.float32)
    R = R.astype(np.float32)
    t = t.astype(np.float32)
    d = []
    for i in range(3):
        x = ((R[i, 0] * q[:, 0] + R[i, 1] * q[:, 1]) + R[i, 2] * q[:, 2]) + t[i]
        d.append(x - p[:, i])
    return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]


TAU_SCHEDULE = (8.0, 4.0, 2.0, 1.0)  # coarse-to-fine residual thresholds, multiples of tau


def masker_spec(points, flow, tau=0.10, sem=None, inst=None, movable=()):
    """Residual-vs-rigid-flow dynamic mask with per-instance voting + final static-point pose.

    1. w = 1 (noSeg) or [sem not in movable] (Seg; all-ones when fewer than 3 such points)
    2. for m in TAU_SCHEDULE:  (R0,t0) = kabsch(p+f -> p; w);  r2_i = ||R0(p_i+f_i)+t0 - p_i||^2 (fp32);
       dyn_i = r2_i > fl32(fl32(tau)*m)^2;  w = 1 - dyn        (a single least-squares fit is biased by
       the movers themselves, so the static set is tightened over four rounds)
    3. after the last round, per instance id >= 1: dynamic iff 2*sum(dyn) > count (strict majority),
       broadcast to members; id 0 keeps its point vote
    4. mask = dyn (u8); final pose = kabsch(p+f -> p; w = 1-mask)
    Returns dict(mask u8[N], bg_index i64, quat wxyz, R, t, R0, t0 (last round), odom f64[7] = [t, x,y,z,w]).
    """
    p = np.asarray(points, np.float32)
    f = np.asarray(flow, np.float32)
    n = p.shape[0]
    w = np.ones(n, np.float64)
    if sem is not None and len(movable):
        cand = (~np.isin(np.asarray(sem), list(movable))).astype(np.float64)
        if cand.sum() >= 3:
            w = cand
    for m in TAU_SCHEDULE:
        _, R0, t0 = kabsch_spec(p, f, w)
        tau_r = np.float32(tau) * np.float32(m)
        dyn = residual_sq(p, f, R0, t0) > tau_r * tau_r
        w = 1.0 - dyn.astype(np.float64)
    if inst is not None:
        inst = np.asarray(inst).astype(np.int64)
        k = int(inst.max()) + 1 if n else 0
        cnt = np.bincount(inst, minlength=k)
        dcnt = np.bincount(inst, weights=dyn.astype(np.int64), minlength=k).astype(np.int64)
        vote = 2 * dcnt > cnt
        has = inst >= 1
        dyn = np.where(has, vote[inst], dyn)
    mask = dyn.astype(np.uint8)
    q, R, t = kabsch_spec(p, f, 1.0 - mask.astype(np.float64))
    return dict(mask=mask, bg_index=np.flatnonzero(mask == 0), quat=q, R=R, t=t, R0=R0, t0=t0,
                odom=np.array([t[0], t[1], t[2], q[1], q[2], q[3], q[0]]))
