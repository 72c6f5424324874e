"""oracle/gmm.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy (float64) restatement of the reference's noSeg masker (scripts/PointCloudOdometry_noSeg.py:97-103,
ASF/main_sju_occ_ros.py:257-263):

    gmm = GaussianMixture(n_components=2).fit_predict(hstack(flow, points))     # sklearn, all defaults
    bg  = Counter(labels).most_common(1)[0][0];  bg_index = argwhere(labels == bg)

The arithmetic lives in a third-party dependency that is not under /root/reference: scikit-learn (requirement.txt:6 unpinned,
:13 `==0.21.3`; this image has 1.9.0).  Its published algorithm (sklearn/mixture/_base.py `fit_predict`,
_gaussian_mixture.py `_estimate_gaussian_parameters`, `_compute_precision_cholesky`, `_estimate_log_gaussian_prob`),
restated here for covariance_type='full', n_components=2, tol=1e-3, reg_covar=1e-6, max_iter=100, n_init=1:

    init: one-hot responsibilities from k-means -> (weights, means, covariances) by the M-step formulas
    loop n_iter = 1..max_iter:
        E: log p_ik = -0.5 (D log 2pi + ||(x_i - mu_k) PC_k||^2) + log det PC_k + log w_k;  lse_i = logsumexp_k;  r_ik = exp(log p_ik - lse_i)
           lower_bound = mean_i lse_i
        M: n_k = sum_i r_ik + 10 eps;  mu_k = sum_i r_ik x_i / n_k;  S_k = sum_i r_ik (x_i-mu_k)(x_i-mu_k)^T / n_k + reg I;
           w_k = n_k / sum n_k;  PC_k = chol(S_k)^-T
        stop when |lower_bound - previous| < tol
    final E step -> labels = argmax_k

The only thing the reference leaves undefined is the k-means seeding (`random_state=None`: a different answer on every run).
The product defines it (SURVEY Appendix C style): features are centred; the two seeds are the points with the smallest and the
largest value of the highest-variance feature (lowest index on ties); Lloyd iterations until no label changes (<= 30).

PARITY: pinned against scikit-learn itself -- `check_against_sklearn` hands sklearn the same initial parameters
(weights_init / means_init / precisions_init) and requires identical labels, identical n_iter and the same lower bound
(oracle/gen_golden_gmm.py -> tests/golden/gmm_mask.npz; tests/test_oracle.py also runs the comparison live).
"""
from collections import Counter

import numpy as np

TOL, REG_COVAR, MAX_ITER, LLOYD_MAX = 1e-3, 1e-6, 100, 30
EPS10 = 10 * np.finfo(np.float64).eps


def features(points, flow):
    """[flow | xyz] as the reference stacks them, float64, centred per column."""
    x = np.concatenate((np.asarray(flow, np.float32), np.asarray(points, np.float32)), axis=1).astype(np.float64)
    return x - x.mean(axis=0)


def kmeans2_init(x):
    """Deterministic 2-means -> labels {0,1} [N]."""
    d = int(np.argmax(x.var(axis=0)))
    c = np.stack([x[int(np.argmin(x[:, d]))], x[int(np.argmax(x[:, d]))]])
    lab = None
    for _ in range(LLOYD_MAX):
        d0 = ((x - c[0]) ** 2).sum(axis=1)
        d1 = ((x - c[1]) ** 2).sum(axis=1)
        new = (d1 < d0).astype(np.int64)          # ties -> cluster 0
        if lab is not None and np.array_equal(new, lab):
            break
        lab = new
        for k in range(2):
            if (lab == k).any():
                c[k] = x[lab == k].mean(axis=0)
    return lab


def m_step(x, resp):
    nk = resp.sum(axis=0) + EPS10
    means = (resp.T @ x) / nk[:, None]
    cov = np.empty((2, x.shape[1], x.shape[1]))
    for k in range(2):
        diff = x - means[k]
        cov[k] = (resp[:, k] * diff.T) @ diff / nk[k]
        cov[k].flat[:: x.shape[1] + 1] += REG_COVAR
    return nk / nk.sum(), means, cov


def precision_cholesky(cov):
    """PC_k = chol(S_k)^-T (upper triangular), as sklearn's _compute_precision_cholesky."""
    pc = np.empty_like(cov)
    for k in range(cov.shape[0]):
        L = np.linalg.cholesky(cov[k])
        pc[k] = np.linalg.solve(L, np.eye(L.shape[0])).T
    return pc


def e_step(x, w, means, pc):
    n, D = x.shape
    logp = np.empty((n, 2))
    for k in range(2):
        y = (x - means[k]) @ pc[k]
        logp[:, k] = -0.5 * (D * np.log(2 * np.pi) + (y * y).sum(axis=1)) + np.log(np.diag(pc[k])).sum() + np.log(w[k])
    mx = logp.max(axis=1)
    lse = mx + np.log(np.exp(logp - mx[:, None]).sum(axis=1))
    return lse.mean(), logp - lse[:, None], logp


def gmm_spec(points, flow):
    """-> dict(labels i64[N], mask u8[N] (1 = not background), bg_index, n_iter, lower_bound, converged, init=(w, means, cov))."""
    x = features(points, flow)
    n = x.shape[0]
    resp = np.zeros((n, 2))
    resp[np.arange(n), kmeans2_init(x)] = 1.0
    w, means, cov = m_step(x, resp)
    init = (w.copy(), means.copy(), cov.copy())
    pc = precision_cholesky(cov)
    lb, converged, n_iter = -np.inf, False, 0
    for n_iter in range(1, MAX_ITER + 1):
        prev = lb
        lb, log_resp, _ = e_step(x, w, means, pc)
        w, means, cov = m_step(x, np.exp(log_resp))
        pc = precision_cholesky(cov)
        if abs(lb - prev) < TOL:
            converged = True
            break
    _, log_resp, logp = e_step(x, w, means, pc)
    labels = (log_resp[:, 1] > log_resp[:, 0]).astype(np.int64)       # argmax, ties -> 0
    c1 = int(labels.sum())
    bg = 1 if c1 > n - c1 else (0 if c1 < n - c1 else int(labels[0]))   # Counter.most_common(1): ties -> first seen
    mask = (labels != bg).astype(np.uint8)
    return dict(labels=labels, mask=mask, bg_index=np.flatnonzero(mask == 0), n_iter=n_iter, lower_bound=lb, converged=converged,
                init=init, margin=np.abs(logp[:, 1] - logp[:, 0]))


def check_against_sklearn(points, flow):
    """Runs scikit-learn's GaussianMixture from the same initial parameters; returns (spec dict, sklearn labels, model)."""
    from sklearn.mixture import GaussianMixture
    spec = gmm_spec(points, flow)
    w0, mu0, cov0 = spec["init"]
    gm = GaussianMixture(n_components=2, weights_init=w0, means_init=mu0, precisions_init=np.linalg.inv(cov0))
    labels = gm.fit_predict(features(points, flow))
    return spec, labels, gm


def reference_bg_index(labels):
    """The reference's two lines after fit_predict."""
    bg = Counter(labels.tolist()).most_common(1)[0][0]
    return np.argwhere(labels == bg).flatten()
