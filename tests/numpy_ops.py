"""Test double: the device primitives of ssf_slam_b200.dataset.subsample_points restated with NumPy, so that the CPU suite can
check the HOST logic (np.random call order, index-list algebra) against the reference golden without a GPU.  Lives in tests/
on purpose -- the product has no CPU path."""
import numpy as np
import torch


class NumpyOps:
    def upload(self, a, dtype):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype)

    def select(self, pts, mask, pre, n, ground_cut, mask_mode):
        src = np.arange(n) if pre is None else pre.numpy()[:n].astype(np.int64)
        keep = np.ones(n, bool)
        if ground_cut:
            keep &= np.logical_not(pts.numpy()[src, -1] < np.float32(-3.3))
        if mask_mode:
            m = mask.numpy()[src]
            keep &= (m != 0) if mask_mode == 1 else ((m == 0) if mask_mode == 2 else (m == 1))
        out = np.zeros(max(n, 1), np.int32)
        k = int(keep.sum())
        out[:k] = src[keep]
        return torch.from_numpy(out), k

    def compose(self, sel, n_sel, ind):
        ind = np.asarray(ind, np.int64)
        assert ((ind >= 0) & (ind < n_sel)).all()
        return torch.from_numpy((ind if sel is None else sel.numpy()[ind]).astype(np.int32))

    def concat(self, a, b):
        return torch.cat([a, b])

    def rows(self, src, lst):
        return src[lst.long()]

    def bytes_(self, src, lst):
        return src[lst.long()]
