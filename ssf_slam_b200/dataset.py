"""CARLA scene-flow frames: the step immediately before the hot path, with the data movement on the device.

Mirrors ``ASF/utils/datasets/carla.py``: ``load_sequence`` (:432-488, npz keys ``pos1 pos2 ego_flow gt [pre_ego_flow
pre_gt] [s_fg_mask t_fg_mask]``) and ``subsample_points`` (:202-305, with ``hybrid_sample_points`` :179-200).

The reference filters, compacts and gathers NumPy arrays on the host.  Here a raw frame is uploaded once (``DeviceFrame``) and
every intermediate cloud of the subsampler is an *index list into the raw arrays* held in HBM: the ground cut and the
foreground / background selections are stable compactions (``ssf_dataset_select``), a random draw is applied by composing index
lists (``ssf_index_compose``), and only the final lists gather point / flow rows and mask bytes (``ssf_gather_rows``,
``ssf_gather_u8``).  What crosses PCIe is the list lengths going down (they size the draws) and the drawn indices going up.
The draws themselves stay on the host and follow the reference's ``np.random`` call sequence exactly, so under the same NumPy
seed the selected points are identical to the reference's (pinned by ``tests/golden/carla_subsample.npz``, written by the
reference's own method -- ``oracle/gen_golden_dataset.py``).

``subsample_points`` takes its device primitives from an ``ops`` object (default: the CUDA kernels, no fallback); the CPU test
suite drives the same host logic with a NumPy stand-in that lives in ``tests/``.
"""
import numpy as np
import torch

from . import _native as nat

GROUND_Z = -3.3   # carla.py:237,243


def load_sequence(source):
    """``source``: path of an .npz or a mapping with the same keys -> (sequence [pos1, pos2], ground_truth, mask) exactly
    as ``CARLA3D.load_sequence`` with the shipped flag ``add_Seg_after_FLow = False``."""
    data = np.load(source) if isinstance(source, (str, bytes)) else source
    try:
        sequence = [np.asarray(data["pos1"]), np.asarray(data["pos2"])]
        if "pre_ego_flow" not in data:
            ground_truth = [np.asarray(data["ego_flow"]), np.asarray(data["gt"])]
        else:
            ground_truth = [np.asarray(data[k]) for k in ("ego_flow", "gt", "pre_ego_flow", "pre_gt")]
        mask = [np.asarray(data["s_fg_mask"]), np.asarray(data["t_fg_mask"])] if ("s_fg_mask" in data and "t_fg_mask" in data) else []
    finally:
        if hasattr(data, "close"):
            data.close()
    return sequence, ground_truth, mask


class CudaOps:
    """The device primitives of the subsampler (csrc/dataset.cu through the C ABI).  Lists are int32 CUDA tensors."""

    def __init__(self, device="cuda:0"):
        nat.require_device(device)
        self.device = torch.device(device)

    def upload(self, a, dtype):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def select(self, pts, mask, pre, n, ground_cut, mask_mode):
        """-> (list of raw indices i32 [n] (first `count` valid), count) ; one 4-byte D2H for the count."""
        sel = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        cnt = torch.empty(1, dtype=torch.int32, device=self.device)
        nat.check(nat.lib().ssf_dataset_select(nat.ptr(pts), 3, nat.ptr(mask), nat.ptr(pre), n, 1 if ground_cut else 0, GROUND_Z,
                                               mask_mode, nat.ptr(sel), nat.ptr(cnt), nat.stream()))
        return sel, int(cnt.item())

    def compose(self, sel, n_sel, ind):
        """out[j] = sel[ind[j]] (sel None = identity); `ind` is a host int array (the np.random draw)."""
        ind_d = self.upload(ind, torch.int32)
        out = torch.empty(ind_d.numel(), dtype=torch.int32, device=self.device)
        err = torch.zeros(1, dtype=torch.int32, device=self.device)
        nat.check(nat.lib().ssf_index_compose(nat.ptr(sel), n_sel, nat.ptr(ind_d), ind_d.numel(), nat.ptr(out), nat.ptr(err), nat.stream()))
        return out

    def concat(self, a, b):
        return torch.cat([a, b])

    def rows(self, src, lst):
        """src f32 [n,C] raw array, lst i32 [m] -> [m,C]"""
        n, C = src.shape
        out = torch.empty(lst.numel(), C, dtype=torch.float32, device=self.device)
        nat.check(nat.lib().ssf_gather_rows(nat.ptr(src), nat.ptr(lst), 1, n, lst.numel(), C, nat.ptr(out), nat.stream()))
        return out

    def bytes_(self, src, lst):
        out = torch.empty(lst.numel(), dtype=torch.uint8, device=self.device)
        nat.check(nat.lib().ssf_gather_u8(nat.ptr(src), src.numel(), nat.ptr(lst), lst.numel(), nat.ptr(out), nat.stream()))
        return out


def _binary_mask(m):
    m = np.asarray(m)
    if not np.isin(m, (0, 1)).all():
        raise ValueError("foreground masks must be binary (0 / 1), as the dataset's s_fg_mask / t_fg_mask")
    return m.astype(np.uint8)


class DeviceFrame:
    """A raw frame (``load_sequence`` output) resident on the device: uploaded once, subsampled any number of times."""

    def __init__(self, sequence, ground_truth, mask, ops=None):
        self.ops = ops if ops is not None else CudaOps()
        up = self.ops.upload
        self.n = [int(sequence[0].shape[0]), int(sequence[1].shape[0])]
        self.pts = [up(np.asarray(sequence[0], np.float32), torch.float32), up(np.asarray(sequence[1], np.float32), torch.float32)]
        self.gt_1d = [np.asarray(g).ndim == 1 for g in ground_truth]
        self.gt = [up(np.asarray(g, np.float32).reshape(len(g), -1), torch.float32) for g in ground_truth]
        self.mask = [up(_binary_mask(m), torch.uint8) for m in mask]


def subsample_points(frame, nb_points, rm_ground=False, hybrid_sample=False, pre_segfrnt=True, rng=np.random):
    """``CARLA3D.subsample_points`` (carla.py:202-305) on a ``DeviceFrame``; flags = the class's attributes (defaults = its
    constructor defaults, :80-107; ``use_fg_inds`` is taken as True, the masks always follow their points).
    -> (sequence [pos1, pos2] f32 [nb,3], ground_truth [.. f32 [nb,C]], mask [u8 [nb], u8 [nb]]) as device tensors.
    ``rng`` defaults to the global NumPy RNG the reference draws from, and is called in the reference's order."""
    ops = frame.ops
    if hybrid_sample and not pre_segfrnt:
        raise ValueError("hybrid_sample without pre_segfrnt leaves the reference's masks and clouds with different lengths")
    if (hybrid_sample or pre_segfrnt) and len(frame.mask) < 2:
        raise ValueError("this flag setting needs the frame's s_fg_mask / t_fg_mask")
    cur = [None, None]                 # index list of each cloud into its raw arrays (None = every raw point, in order)
    cnt = list(frame.n)
    if rm_ground:                      # :236-247
        for c in (0, 1):
            cur[c], cnt[c] = ops.select(frame.pts[c], None, None, frame.n[c], True, 0)
    mcur, mcnt = list(cur), list(cnt)  # the clouds the MASK arrays are aligned with (see below)
    if hybrid_sample:                  # :249-251 with hybrid_sample_points(mask, num_pts=100), :179-200
        for c in (0, 1):
            fg, n_fg = ops.select(None, frame.mask[c], cur[c], cnt[c], False, 3)      # positions where mask == 1
            bg, n_bg = ops.select(None, frame.mask[c], cur[c], cnt[c], False, 2)      # positions where mask == 0
            if n_fg < 100:
                # every foreground point, then background draws -- which the reference applies as positions of the CURRENT
                # cloud, not of its background list (:187-190); kept, the golden pins it
                draw = rng.choice(n_bg, nb_points - n_fg, replace=False)
                cur[c] = ops.concat(fg[:n_fg], ops.compose(cur[c], cnt[c], draw))
            else:
                draw_fg = rng.choice(n_fg, 100, replace=False)
                draw_bg = rng.choice(n_bg, nb_points - 100, replace=False)
                cur[c] = ops.concat(ops.compose(fg, n_fg, draw_fg), ops.compose(bg, n_bg, draw_bg))
            cnt[c] = int(cur[c].shape[0])
        mcur, mcnt = list(cur), list(cnt)      # hybrid sampling also replaces the masks (mask[src_ind], :199)
    elif pre_segfrnt:                  # :262-265: foreground points only -- the reference compacts the clouds and the flows but
        for c in (0, 1):               # NOT the masks, which are then indexed with the compacted clouds' draws (:299-300); kept
            cur[c], cnt[c] = ops.select(None, frame.mask[c], cur[c], cnt[c], False, 1)
    final, mfinal = [], []
    for c in (0, 1):                   # :270-282
        draw = rng.choice(cnt[c], nb_points, replace=not (nb_points < cnt[c]))
        final.append(ops.compose(cur[c], cnt[c], draw))
        mfinal.append(final[c] if mcur[c] is cur[c] else ops.compose(mcur[c], mcnt[c], draw))
    sequence = [ops.rows(frame.pts[0], final[0]), ops.rows(frame.pts[1], final[1])]
    if frame.gt_1d[0]:                 # :285-286: only the first array follows the draw
        ground_truth = [ops.rows(frame.gt[0], final[0]).reshape(-1)] + [
            (ops.rows(g, cur[0][:cnt[0]]) if cur[0] is not None else g) for g in frame.gt[1:]]
        mask = [ops.bytes_(m, mcur[c][:mcnt[c]]) if mcur[c] is not None else m for c, m in enumerate(frame.mask)]
    else:
        ground_truth = [ops.rows(g, final[0]) for g in frame.gt]
        mask = [ops.bytes_(frame.mask[0], mfinal[0]), ops.bytes_(frame.mask[1], mfinal[1])] if len(frame.mask) >= 2 else []
    return sequence, ground_truth, mask


def device_batch(frames):
    """list of ``subsample_points`` results with equal point counts -> dict of device tensors: pos1, pos2 f32 [B,N,3],
    gt f32 [B,N,3] and, when present, s_fg_mask u8 [B,N] (1 = dynamic object, as ``data['mask'][0]`` in the drivers)."""
    out = dict(pos1=torch.stack([f[0][0] for f in frames]), pos2=torch.stack([f[0][1] for f in frames]),
               gt=torch.stack([f[1][1][:, :3] for f in frames]).contiguous())
    if all(len(f[2]) >= 1 for f in frames):
        out["s_fg_mask"] = torch.stack([f[2][0] for f in frames])
    return out
