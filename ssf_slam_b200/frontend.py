"""Host-facing front end: the calls the reference's ROS drivers make after (or instead of) the network.

Mirrors, with NumPy arrays in and out exactly like the reference's in-process code:

* ``slove_RT_by_SVD(src, dst) -> (R[3,3], t[3,1])``  -- scripts/PointCloudOdometry.py:15-33 (same spelling);
  computed on the GPU by the Kabsch reduction kernel (fp64 accumulators, Horn closed form) instead of
  ``np.linalg.svd``.  The reflection branch returns the proper rotation the reference intended
  (its ``Vt.T & U.T`` raises TypeError -- deliberate divergence).
* ``background_index(points, flow, ...) -> bg_index`` -- stands where the drivers compute ``bg_index`` from the GMM
  (ASF/main_sju_occ_ros.py:257-263) or the GT mask (scripts/PointCloudOdometry.py:91): the deterministic
  residual-vs-rigid-flow masker with optional per-instance voting.
* ``gmm_background(points, flow) -> bg_index`` -- the reference's own noSeg masker, the three lines
  ``GaussianMixture(n_components=2).fit_predict(hstack(flow, points))`` / ``Counter.most_common`` / ``argwhere``
  (scripts/PointCloudOdometry_noSeg.py:97-103, ASF/main_sju_occ_ros.py:257-263), fitted on the GPU by EM with scikit-learn's
  defaults and a deterministic seeding (csrc/gmm.cu; specification pinned against scikit-learn in oracle/gmm.py).
* ``odometry(points, flow, ...)`` -- mask + pose + the ``frame_odom1`` payload ``[tx,ty,tz,qx,qy,qz,qw]`` in one launch.
* ``SceneFlowFrontEnd`` -- network + mask + ego-motion for batches of frame pairs held in HOST memory (the end-to-end
  call ``bench.py`` times: H2D of the clouds, all kernels, D2H of masks and poses).
"""
import warnings

import numpy as np
import torch

from . import functional as F_
from . import _native as nat


def _dev(a, dtype, device):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device)


def slove_RT_by_SVD(src, dst, device="cuda:0"):
    """Un-weighted rigid fit dst ~= R @ src + t over all rows; src, dst [M,3] array-likes -> (R [3,3], t [3,1]) float64.
    Precision follows the input like the reference's numpy code does: float64 clouds are reduced in float64 end to end
    (`ssf_solve_rt_f64`; R within 1e-9 and t within 1e-6 of the reference), float32 clouds are read as float32 and accumulated in
    float64.  Fewer than 3 points cannot fix a rotation: the identity is returned with a warning (numpy's SVD would return an
    arbitrary member of the solution set)."""
    nat.require_device(device)
    src, dst = np.asarray(src), np.asarray(dst)
    if src.shape != dst.shape or src.ndim != 2 or src.shape[1] != 3:
        raise ValueError("slove_RT_by_SVD: src and dst must both be [M,3]")
    if src.shape[0] < 3:
        warnings.warn("slove_RT_by_SVD: fewer than 3 point pairs, returning the identity pose")
    if src.dtype == np.float32 and dst.dtype == np.float32:
        src_t = torch.from_numpy(np.ascontiguousarray(src)).to(device).unsqueeze(0)
        dst_t = torch.from_numpy(np.ascontiguousarray(dst)).to(device).unsqueeze(0)
        _, _, pose = F_.frontend(dst_t, src_t, mode=2, want_pose=True)
    else:
        src_t = torch.from_numpy(np.ascontiguousarray(src, np.float64)).to(device).unsqueeze(0)
        dst_t = torch.from_numpy(np.ascontiguousarray(dst, np.float64)).to(device).unsqueeze(0)
        _, pose = F_.solve_rt_f64(src_t, dst_t)
    pose = pose[0].cpu().numpy()
    return pose[:9].reshape(3, 3).copy(), pose[9:].reshape(3, 1).copy()


def odometry(points, flow, mask=None, sem=None, inst=None, movable=(), tau=0.10, device="cuda:0", masker="residual"):
    """points, flow [N,3] (or [B,N,3]) -> dict(mask u8, bg_index, odom f64[7] = [t, qx,qy,qz,qw], R, t).
    ``mask`` given (0 = background, as ``s_fg_mask``) -> pose from those points only (GT-mask variants);
    otherwise the residual masker runs (noSeg), seeded/voted by ``sem``/``inst`` when given (Seg), or -- ``masker="gmm"`` --
    the reference's 2-component Gaussian-mixture masker."""
    nat.require_device()
    p = np.asarray(points, np.float32)
    single = p.ndim == 2
    if single:
        p = p[None]
    f = np.asarray(flow, np.float32).reshape(p.shape)
    tp, tf = _dev(p, torch.float32, device), _dev(f, torch.float32, device)
    if mask is not None:
        tm = _dev(np.asarray(mask).reshape(p.shape[:2]) != 0, torch.uint8, device)
        m, odom, pose = F_.frontend(tp, tf, mode=0, in_mask=tm, want_pose=True)
    elif masker == "gmm":
        tm = F_.gmm_mask(tp, tf)
        m, odom, pose = F_.frontend(tp, tf, mode=0, in_mask=tm, want_pose=True)
    else:
        ts = None if sem is None else _dev(np.asarray(sem).reshape(p.shape[:2]), torch.int32, device)
        ti = None if inst is None else _dev(np.asarray(inst).reshape(p.shape[:2]), torch.int32, device)
        n_inst = 0 if inst is None else int(np.max(inst)) + 1
        m, odom, pose = F_.frontend(tp, tf, mode=1, sem=ts, movable=movable, inst=ti, n_inst=n_inst, tau=tau, want_pose=True)
    m, odom, pose = m.cpu().numpy(), odom.cpu().numpy(), pose.cpu().numpy()
    out = dict(mask=m, odom=odom, R=pose[:, :9].reshape(-1, 3, 3), t=pose[:, 9:])
    if single:
        out = {k: v[0] for k, v in out.items()}
        out["bg_index"] = np.flatnonzero(out["mask"] == 0)
    else:
        out["bg_index"] = [np.flatnonzero(mk == 0) for mk in out["mask"]]   # ragged: one ascending index array per cloud
    return out


def background_index(points, flow, sem=None, inst=None, movable=(), tau=0.10, device="cuda:0"):
    """bg_index (ascending int64) of the static points, the quantity the drivers feed to slove_RT_by_SVD; for batched
    [B,N,3] input a list of B such arrays."""
    return odometry(points, flow, sem=sem, inst=inst, movable=movable, tau=tau, device=device)["bg_index"]


def gmm_background(points, flow, device="cuda:0"):
    """bg_index of the majority component of a 2-component GMM on [flow | xyz]: what the reference's noSeg drivers compute
    with scikit-learn on the host (scripts/PointCloudOdometry_noSeg.py:97-103)."""
    return odometry(points, flow, device=device, masker="gmm")["bg_index"]


class _Pending:
    """Result of ``SceneFlowFrontEnd.submit``: pinned host buffers that are valid once the recorded event has completed."""

    def __init__(self, event, out):
        self._event, self._out = event, out

    def result(self):
        self._event.synchronize()
        return self._out


class SceneFlowFrontEnd:
    """Scene flow + dynamic mask + ego-motion for batches of frame pairs in host memory.

    ``process`` is the synchronous call.  ``submit(..., slot=i)`` runs the same work on the i-th of ``n_slots`` private CUDA
    streams with private pinned staging buffers and returns immediately; independent batches submitted to different slots
    overlap on the GPU (one batch's H2D / D2H copies and its latency-bound kernels such as FPS run under the other's dense
    kernels).  Frame pairs are independent (ASF/main_sju_occ_ros.py:168-284), so this is plain pipelining."""

    def __init__(self, net, device="cuda:0", tau=0.10, movable=(), n_slots=2, use_graph=False, masker="residual"):
        nat.require_device()
        if masker not in ("residual", "gmm"):
            raise ValueError("masker must be 'residual' (DESIGN.md section 5) or 'gmm' (the reference's noSeg masker)")
        self.net, self.device, self.tau, self.movable, self.masker = net, torch.device(device), tau, tuple(movable), masker
        self._pin = {}
        self._streams = [torch.cuda.Stream(device=self.device) for _ in range(n_slots)]
        self._pending = [None] * n_slots
        self.use_graph = use_graph
        self._graphs = {}   # (slot, B, N, seg, return_flow) -> (graph, static inputs, static outputs)
        if hasattr(net, "weights"):   # prepare the kernel-side weight images once, before any slot stream can race on them
            net.weights(self.device)
            torch.cuda.synchronize(self.device)

    def _staged(self, name, arr, dtype, out=None):
        """host array -> pinned staging buffer -> device (async on the current stream; into `out` when given)."""
        t = torch.as_tensor(arr)
        key = (name, tuple(t.shape), dtype)
        if key not in self._pin:
            self._pin[key] = torch.empty(t.shape, dtype=dtype, pin_memory=True)
        self._pin[key].copy_(t)
        if out is not None:
            out.copy_(self._pin[key], non_blocking=True)
            return out
        return self._pin[key].to(self.device, non_blocking=True)

    def _host_out(self, name, t):
        key = (name, tuple(t.shape), t.dtype)
        if key not in self._pin:
            self._pin[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        self._pin[key].copy_(t, non_blocking=True)
        return self._pin[key]

    def _run(self, x1, x2, ts, ti, n_inst):
        flows, _ = self.net.forward_pm(x1, x2)
        flow = flows[0] if flows[0].shape[-1] == 3 else flows[0][..., :3].contiguous()   # 4-channel heads: xyz flow only
        if self.masker == "gmm":
            mask, odom = F_.frontend(x1, flow, mode=0, in_mask=F_.gmm_mask(x1, flow))
        else:
            mask, odom = F_.frontend(x1, flow, mode=1, sem=ts, movable=self.movable, inst=ti, n_inst=n_inst, tau=self.tau)
        return flows[0], mask, odom

    def _graph_for(self, slot, B, N, seg, n_inst):
        """CUDA graph of the whole step (172 kernel launches at N = 8192) for one slot and input shape: static device
        inputs / outputs, captured after two eager warm-up runs on the slot's stream.  Replaying it removes the per-launch
        host cost, which dominates single-pair latency (the reference's operating point: one pair per 100 ms)."""
        key = (slot, B, N, seg, n_inst)
        if key not in self._graphs:
            dev = self.device
            x1 = torch.zeros(B, N, 3, dtype=torch.float32, device=dev)
            x2 = torch.zeros(B, N, 3, dtype=torch.float32, device=dev)
            ts = torch.zeros(B, N, dtype=torch.int32, device=dev) if seg else None
            ti = torch.zeros(B, N, dtype=torch.int32, device=dev) if seg else None
            # warm-up on well-formed clouds (distinct points) so every lazily initialised path has run before capture
            g = torch.Generator(device=dev).manual_seed(0)
            x1.copy_(torch.randn(B, N, 3, device=dev, generator=g) * 20)
            x2.copy_(x1 + 0.1)
            for _ in range(2):
                self._run(x1, x2, ts, ti, n_inst)
            torch.cuda.current_stream(dev).synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=torch.cuda.current_stream(dev)):
                flow, mask, odom = self._run(x1, x2, ts, ti, n_inst)
            self._graphs[key] = (graph, (x1, x2, ts, ti), (flow, mask, odom))
        return self._graphs[key]

    @torch.no_grad()
    def submit(self, pos1, pos2, sem=None, inst=None, n_inst=0, return_flow=False, slot=0):
        """Asynchronous ``process`` on slot ``slot``; returns a handle whose ``.result()`` yields the same dict.  The
        buffers of a slot are reused by its next submit, which first waits for the previous one."""
        if self._pending[slot] is not None:
            self._pending[slot].result()   # staging buffers of this slot are free again
        st = self._streams[slot]
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st):
            tag = "s%d." % slot
            if self.use_graph:
                B, N = int(np.shape(pos1)[0]), int(np.shape(pos1)[1])
                graph, (g1, g2, gs, gi), (flow, mask, odom) = self._graph_for(slot, B, N, sem is not None, int(n_inst))
                self._staged(tag + "p1", pos1, torch.float32, out=g1)
                self._staged(tag + "p2", pos2, torch.float32, out=g2)
                if sem is not None:
                    self._staged(tag + "sem", sem, torch.int32, out=gs)
                    self._staged(tag + "inst", inst, torch.int32, out=gi)
                graph.replay()
            else:
                x1 = self._staged(tag + "p1", pos1, torch.float32)
                x2 = self._staged(tag + "p2", pos2, torch.float32)
                ts = None if sem is None else self._staged(tag + "sem", sem, torch.int32)
                ti = None if inst is None else self._staged(tag + "inst", inst, torch.int32)
                flow, mask, odom = self._run(x1, x2, ts, ti, n_inst)
            out = dict(mask=self._host_out(tag + "mask", mask), odom=self._host_out(tag + "odom", odom))
            if return_flow:
                out["flow"] = self._host_out(tag + "flow", flow)
            ev = torch.cuda.Event()
            ev.record(st)
        self._pending[slot] = _Pending(ev, out)
        return self._pending[slot]

    def process(self, pos1, pos2, sem=None, inst=None, n_inst=0, return_flow=False):
        """pos1, pos2: host f32 [B,N,3] -> dict(mask u8 [B,N], odom f64 [B,7] (, flow f32 [B,N,3])) on the host
        (pinned buffers owned by this object, overwritten by the next call on the same slot)."""
        return self.submit(pos1, pos2, sem=sem, inst=inst, n_inst=n_inst, return_flow=return_flow, slot=0).result()

    def h2d_bytes(self, B, N, seg=False):
        return B * N * 3 * 4 * 2 + (B * N * 4 * 2 if seg else 0)

    def d2h_bytes(self, B, N, return_flow=False):
        return B * N + B * 7 * 8 + (B * N * 12 if return_flow else 0)
