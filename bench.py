#!/usr/bin/env python
"""bench.py -- scene-flow frame-pairs/s @N=8192 (BASELINE.json metric) for the B200-native front end.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference] [--masker residual|gmm] [--config 1|2|4]

A *step* is one pass of the hot path (scene-flow network + dynamic mask + ego-motion) over one batch of B synthetic
CARLA-shaped frame pairs drawn from a 200-frame sequence (BASELINE.json configs[1]; weak scaling: every rank processes its own
sequences, configs[3]); consecutive steps take consecutive windows of the sequence, so no two timed steps see the same batch.
`value` = frame pairs per second over all ranks with inputs resident in HBM; `e2e` = the same through
`SceneFlowFrontEnd.submit` from HOST buffers (pinned staging, H2D + kernels + D2H of masks and poses inside the timed region).
`--impl reference` times the reference pipeline's CPU implementation (the oracle port of the reference's Python: TFlow port
pinned bit-exact to the unmodified reference + the SAME masker as the GPU arm + slove_RT_by_SVD) on the host cores.
`--config 2` = BASELINE configs[2] (Seg pipeline, N = 16384, semantic seed + per-instance voting), `--config 4` = configs[4]
(N = 65536 FPS / kNN / ball-query radius sweep); the default and the driver's line is configs[1].
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "frame-pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4], help="index into BASELINE.json configs: 1 (default; 3 is "
                    "the same workload sharded over --gpus ranks), 2 = Seg pipeline N=16384, 4 = N=65536 point-op stress sweep")
    ap.add_argument("--batch", type=int, default=None, help="frame pairs per step per GPU (default 128; 16 for --config 2)")
    ap.add_argument("--npoints", type=int, default=None)
    ap.add_argument("--pool", type=int, default=200, help="distinct synthetic frame pairs per rank (one sequence)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=16, help="frame pairs in the cpu_baseline sample (~0.6 s each on 16 cores)")
    ap.add_argument("--ref-pairs-per-step", type=int, default=2, help="--impl reference: frame pairs per step (bounded sample)")
    ap.add_argument("--graph", action="store_true", help="e2e / latency legs: CUDA-graph replay of the step instead of eager "
                    "launches (measured equal on B200: the step is bound by kernel time, not by launch cost)")
    ap.add_argument("--streams", type=int, default=2, help="CUDA streams independent batches are pipelined over (measured on B200: "
                    "2 streams x 128 pairs 2235 frame-pairs/s, 3 x 64 2182, 2 x 192 2240, 2 x 256 2223)")
    ap.add_argument("--masker", default="residual", choices=["residual", "gmm"], help="dynamic-point masker of the step IN BOTH "
                    "ARMS: the residual-vs-rigid-flow masker (north star, DESIGN.md section 5) or the reference's noSeg GMM masker")
    ap.add_argument("--shares-out", default=None, help="write the full per-kernel CUDA-event table (JSON) to this path")
    ap.add_argument("--no-extras", action="store_true", help="skip the point-op / masker / latency legs (profiling runs)")
    a = ap.parse_args()
    if a.config == 3:
        a.config = 1
    if a.npoints is None:
        a.npoints = {1: 8192, 2: 16384, 4: 65536}[a.config]
    if a.batch is None:
        a.batch = {1: 128, 2: 16, 4: 4}[a.config]
    if a.config == 2:
        a.pool = min(a.pool, 64)
    return a


def metric_name(args):
    if args.config == 4:
        return "point-op stress clouds/s @N=%d (FPS 2048 + kNN16 NxN + ball-query sweep)" % args.npoints
    return "scene-flow frame-pairs/s @N=%d" % args.npoints


def workload_name(args):
    if args.config == 2:
        return ("configs[2]: Seg_ActiveSceneFlow with synthetic semantic/instance labels, per-instance dynamic voting, N=%d "
                "(flow + Seg masker + ego-motion)" % args.npoints)
    if args.config == 4:
        return "configs[4]: dense full-beam stress, N=%d per cloud: FPS / kNN / ball-query radius sweep 0.5-4 m" % args.npoints
    return ("configs[1]: 200-frame synthetic sequence shaped like rm_road/SF/00, N=%d, noSeg_ActiveSceneFlow pipeline "
            "(flow + dynamic mask + ego-motion)" % args.npoints)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, tensor=1590.0, source="fallback")


def traffic_table():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except (OSError, ValueError):
        return {}


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_pipeline_seconds(pool, n_pairs, sd, masker="residual", seg=False, start=0):
    """Reference pipeline on the host, one frame pair at a time as the reference's drivers do: oracle TFlow port (pinned
    bit-exact to the unmodified reference) + the masker of the run (residual: oracle/frontend.py::masker_spec, the NumPy
    statement of the GPU arm's masker; gmm: the reference's own sklearn GMM, majority = background + slove_RT_by_SVD) + the
    frame_odom1 message.  Returns the list of seconds per frame pair."""
    import torch
    from oracle import frontend as ofe
    from oracle import tflow_port as tp
    from ssf_slam_b200 import synth
    torch.set_num_threads(os.cpu_count())
    times = []
    for i in range(n_pairs):
        it = pool[(start + i) % len(pool)]
        t0 = time.perf_counter()
        pc1 = torch.from_numpy(it["pos1"].T.copy()).unsqueeze(0)
        pc2 = torch.from_numpy(it["pos2"].T.copy()).unsqueeze(0)
        flows, _ = tp.tflow_forward(sd, pc1, pc2)
        flow = flows[0][0].numpy().T.copy()
        if masker == "gmm":
            bg = ofe.gmm_background(it["pos1"], flow, random_state=0)
            R, t = ofe.reference_pose(it["pos1"], flow, bg)
            ofe.odom_message(R, t)
        elif seg:
            ofe.masker_spec(it["pos1"], flow, 0.10, sem=it["sem"], inst=it["inst"], movable=synth.MOVABLE_CLASSES)
        else:
            ofe.masker_spec(it["pos1"], flow, 0.10)
        times.append(time.perf_counter() - t0)
    return times


def cpu_stress_seconds(n_points, n_clouds):
    """configs[4] on the host with the plain-C oracle (OpenMP over queries): FPS 2048, kNN16 on a 1/64 query subsample scaled to
    the full N x N search, ball query sweep around the FPS centres.  Returns seconds per cloud (kNN scaled)."""
    from oracle import point_ops as po
    from ssf_slam_b200 import synth
    out = []
    for c in range(n_clouds):
        xyz = synth.dense_cloud(5000 + c, n_points)[None]
        t0 = time.perf_counter()
        fps = po.c_fps(xyz, 2048)
        t_fps = time.perf_counter() - t0
        sub = np.ascontiguousarray(xyz[:, ::64])
        t0 = time.perf_counter()
        po.c_knn(16, sub, xyz)
        t_knn = (time.perf_counter() - t0) * 64.0
        cent = np.ascontiguousarray(xyz[0][fps[0].astype(np.int64)][None])
        t0 = time.perf_counter()
        for r in (0.5, 1.0, 2.0, 4.0):
            po.c_ball_query(r, 16, xyz, cent)
        out.append(t_fps + t_knn + time.perf_counter() - t0)
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    import torch
    from ssf_slam_b200 import synth
    from ssf_slam_b200.weights import random_init_state_dict
    cores = os.cpu_count()
    if args.config == 4:
        cpu_stress_seconds(args.npoints, 1)
        per = [float(sum(cpu_stress_seconds(args.npoints, 1))) for _ in range(args.steps)]
        total, units = float(sum(per)), args.steps
        sample = ("%d clouds of %d points, plain-C oracle with OpenMP: FPS 2048 + kNN16 on every 64th query (x64) + 4-radius ball "
                  "query" % (args.steps, args.npoints))
        unit = "clouds/s"
    else:
        ppstep = max(1, args.ref_pairs_per_step)
        need = ppstep * args.steps + 1
        pool = synth.make_sequence(1000, min(args.pool, need), args.npoints)
        sd = random_init_state_dict(0)
        seg = args.config == 2
        cpu_pipeline_seconds(pool, max(1, min(args.warmup, 1)), sd, args.masker, seg)  # warm-up (one pair; a CPU pair takes ~0.6 s)
        t = cpu_pipeline_seconds(pool, ppstep * args.steps, sd, args.masker, seg, start=1)
        total, units = float(sum(t)), len(t)
        cores = torch.get_num_threads()
        sample = ("%d steps x %d frame pairs (N=%d), one pair at a time: oracle TFlow port (bit-exact to the unmodified reference "
                  "on CPU, C/OpenMP FPS+kNN) + %s" % (args.steps, ppstep, args.npoints,
                                                     "sklearn GMM mask + slove_RT_by_SVD" if args.masker == "gmm" else
                                                     "residual masker spec (NumPy) incl. pose"))
        unit = UNIT
    v = units / total
    cfg = {"workload": workload_name(args) + "; bounded sample on host cores", "masker": args.masker}
    line = {"impl": "reference", "metric": metric_name(args), "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def cost_volume_flops(n1, m):
    """2*MAC of one ssf_cost_volume launch per cloud (algorithmic, after the per-point split of the first layers)."""
    per_row = 2 * m * m + 2 * (m * m + 3 * m + m * m) + 2 * (m * m + m * (m // 2) + m // 2)  # L2 x2, mlp3 x2, weightnet x2
    per_point = 16 * per_row + 16 * 16 * m + 2 * 16 * 16 * m + 16 * m  # + QK^T, two mixes, forward cost
    return 2.0 * n1 * per_point


def _timed(fn, reps=5, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def point_op_rooflines(B, N, dev):
    """Stand-alone pointnet2 operators (the B-op boundary, reference layouts) on B clouds of N points: achieved HBM GB/s =
    ALGORITHMIC bytes / CUDA-event time for the gathers, work rates for FPS / kNN / ball query (not HBM
    bound).  Inputs are larger than L2 in aggregate or freshly produced; 3 warm-ups, 5 timed launches each.  `traffic` = DRAM
    bytes per launch from the committed ncu capture of the same shape (profiles/ncu_traffic.json), when there is one."""
    import torch
    from ssf_slam_b200 import pointnet2_utils as pu
    pk = peaks()
    tr = traffic_table()
    g = torch.Generator(device=dev).manual_seed(0)
    xyz = torch.randn(B, N, 3, device=dev, generator=g) * torch.tensor([30.0, 20.0, 2.0], device=dev)
    M = 2048
    fps_idx = pu.furthest_point_sample(xyz, M)
    new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
    _, idx16 = pu.knn(16, xyz, xyz)
    feat96 = torch.randn(B, 96, N, device=dev, generator=g)
    w3 = torch.rand(B, N, 3, device=dev, generator=g)
    idx3 = idx16[:, :, :3].contiguous()
    out = []

    def traffic(name):
        t = tr.get(name)   # captured at t["batch"] clouds; these operators' traffic is linear in the number of clouds
        return (t["dram_read_bytes"] + t["dram_write_bytes"]) * B / t["batch"] if t is not None and t.get("batch") else None

    def hbm(name, byts, dram, fn):
        """byts = algorithmic bytes (gathered reads counted as memory reads); dram = compulsory DRAM bytes
        (unique source + indices + output): the gathers hit shared memory / L2, so `frac` can exceed 1 while
        `dram_frac` <= 1 is the fraction of the HBM roofline the compulsory stream reaches."""
        t = _timed(fn)
        out.append({"op": name, "bound": "hbm", "achieved": byts / t / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": byts / t / 1e9 / pk["hbm"], "dram_achieved": dram / t / 1e9, "dram_frac": dram / t / 1e9 / pk["hbm"],
                    "ms": t * 1e3, "traffic": traffic(name)})

    def rate(name, work, unit, fn):
        t = _timed(fn)
        out.append({"op": name, "bound": "alu/latency", "achieved": work / t / 1e9, "unit": unit, "ms": t * 1e3, "traffic": traffic(name)})

    C, S = 96, 16
    hbm("grouping_operation[C=96,M=%d,S=16]" % N, B * (4.0 * N * S + 8.0 * C * N * S), B * (4.0 * N * S + 4.0 * C * N + 4.0 * C * N * S),
        lambda: pu.grouping_operation(feat96, idx16))
    hbm("three_interpolate[C=96,n=%d]" % N, B * (4.0 * 3 * N * 2 + 4.0 * C * N * 3 + 4.0 * C * N), B * (4.0 * 3 * N * 2 + 8.0 * C * N),
        lambda: pu.three_interpolate(feat96, idx3, w3))
    # gather_operation: compulsory DRAM reads are whole 32-byte sectors: min(32 M, 4 N) bytes per channel row (with M = N / 4 FPS
    # picks spread over the row nearly every sector holds a picked element, so the row is read once: ncu measures 202 MB read for
    # these 64 clouds = 64 x 96 x 8192 x 4, profiles/ncu_traffic.json); `byts` stays the algorithmic 4 M + 8 C M of SURVEY 8(d)
    hbm("gather_operation[C=96,M=%d]" % M, B * (4.0 * M + 8.0 * C * M), B * (4.0 * M + min(32.0 * M, 4.0 * N) * C + 4.0 * C * M),
        lambda: pu.gather_operation(feat96, fps_idx))
    rate("furthest_point_sample[N=%d,n=%d]" % (N, M), B * float(N) * M, "G point-updates/s", lambda: pu.furthest_point_sample(xyz, M))
    rate("knn[k=16,%dx%d]" % (N, N), B * float(N) * N, "G pair-evaluations/s (brute-force equivalent)", lambda: pu.knn(16, xyz, xyz))
    rate("ball_query[r=1.0,ns=16,%dx%d]" % (M, N), B * float(N) * M, "G pair-evaluations/s", lambda: pu.ball_query(1.0, 16, xyz, new_xyz))
    return out


def run_stress(args, rank, world, local):
    """configs[4]: N = 65536 per cloud, FPS 65536 -> 2048, kNN k = 16 of the cloud in itself, ball query around the FPS centres
    for r in 0.5 / 1 / 2 / 4 m (nsample 16 and 32).  A step = all of it over B clouds; value = clouds/s."""
    import torch
    from ssf_slam_b200 import _native as nat
    from ssf_slam_b200 import pointnet2_utils as pu
    from ssf_slam_b200 import synth
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nat.require_device()
    B, N = args.batch, args.npoints
    host = np.stack([synth.dense_cloud(5000 + rank * B + c, N) for c in range(B)])
    xyz = torch.from_numpy(host).to(dev)
    res = {}

    def step(x):
        fps = pu.furthest_point_sample(x, 2048)
        cent = pu.gather_operation(x.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
        index = pu.build_index(x)            # one spatial index per cloud, shared by the kNN and the eight ball queries
        _, idx = pu.knn(16, x, x, index=index)
        out = [idx]
        for ns in (16, 32):
            for r in (0.5, 1.0, 2.0, 4.0):
                out.append(pu.ball_query(r, ns, x, cent, index=index))
        return fps, out

    for _ in range(max(args.warmup, 1)):
        step(xyz)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(xyz)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = nat.launch_count() - l0
    clk = clocks.stop()
    # end to end: host cloud in (pinned), FPS indices + kNN indices + 8 ball-query index sets back
    pin = torch.from_numpy(host).pin_memory()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    d2h = 0
    for _ in range(args.steps):
        x = pin.to(dev, non_blocking=True)
        fps, outs = step(x)
        hs = [fps.cpu()] + [o.cpu() for o in outs]
        d2h = sum(h.numel() * h.element_size() for h in hs)
    t1.record()
    torch.cuda.synchronize()
    ems = t0.elapsed_time(t1)
    fps_idx = pu.furthest_point_sample(xyz, 2048)
    cent = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
    t = _timed(lambda: pu.furthest_point_sample(xyz, 2048), reps=3, warm=1)
    res["fps_npoint2048"] = {"ms": t * 1e3, "G_point_updates_per_s": B * N * 2048.0 / t / 1e9}
    t = _timed(lambda: pu.build_index(xyz), reps=3, warm=1)
    res["build_index"] = {"ms": t * 1e3}
    index = pu.build_index(xyz)
    t = _timed(lambda: pu.knn(16, xyz, xyz, index=index), reps=3, warm=1)
    res["knn_k16_NxN_indexed"] = {"ms": t * 1e3, "G_pair_evals_per_s_bruteforce_equivalent": B * float(N) * N / t / 1e9}
    t = _timed(lambda: pu.knn(16, xyz, xyz), reps=3, warm=1)
    res["knn_k16_NxN"] = {"ms": t * 1e3, "G_pair_evals_per_s_bruteforce_equivalent": B * float(N) * N / t / 1e9,
                          "GB_per_s_compulsory": B * (24.0 * N + 4.0 * N * 16) / t / 1e9}
    t = _timed(lambda: pu.knn(16, cent, xyz), reps=3, warm=1)
    res["knn_k16_2048xN"] = {"ms": t * 1e3, "G_pair_evals_per_s": B * 2048.0 * N / t / 1e9}
    for ns in (16, 32):
        for r in (0.5, 1.0, 2.0, 4.0):
            t = _timed(lambda: pu.ball_query(r, ns, xyz, cent, index=index), reps=3, warm=1)
            t0 = _timed(lambda: pu.ball_query(r, ns, xyz, cent, use_index=False), reps=3, warm=1)
            _, cnt = pu.ball_query(r, ns, xyz, cent, return_count=True, index=index)
            res["ball_query_r%g_ns%d" % (r, ns)] = {"ms": t * 1e3, "ms_scan": t0 * 1e3, "G_pair_evals_per_s_bruteforce_equivalent": B * 2048.0 * N / t / 1e9,
                                                  "mean_in_range": float(cnt.float().mean().item())}
    line = {"metric": metric_name(args), "value": world * B * args.steps / (ms * 1e-3), "unit": "clouds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic uniform-density full-beam clouds (ssf_slam_b200/synth.py::dense_cloud)",
            "config": {"workload": workload_name(args), "clouds_per_step_per_gpu": B,
                       "l2": "per-step index outputs (~%.0f MB) exceed nothing: these operators are ALU / latency bound, not HBM bound" % (B * N * 16 * 4 / 1e6)},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": world * B * args.steps / (ems * 1e-3), "unit": "clouds/s", "h2d_bytes_per_step": int(host.nbytes),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ems / args.steps},
            "roofline": None, "ops": res, "cpu_baseline": None}
    if not args.no_cpu_baseline and world == 1:
        cpu_stress_seconds(N, 1)
        tt = cpu_stress_seconds(N, 2)
        line["cpu_baseline"] = {"value": len(tt) / float(sum(tt)), "unit": "clouds/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "2 clouds: plain-C oracle (OpenMP): FPS 2048 + kNN16 on every 64th query (x64) + 4-radius ball query"}
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from ssf_slam_b200 import _native as nat
    from ssf_slam_b200 import functional as F_
    from ssf_slam_b200 import profile as prof
    from ssf_slam_b200 import synth
    from ssf_slam_b200.frontend import SceneFlowFrontEnd
    from ssf_slam_b200.model import TFlow
    from ssf_slam_b200.shard import ResultGatherer, my_sequences
    from ssf_slam_b200.weights import random_init_state_dict

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; rank 0's stdout must carry exactly one JSON
        # line, so file descriptor 1 points at stderr until the first collective has completed
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    nat.require_device()
    if args.config == 4:
        run_stress(args, rank, world, local)
        if world > 1:
            dist.destroy_process_group()
        return
    B, N, K, Wm = args.batch, args.npoints, args.steps, args.warmup
    POOL = max(args.pool, 1)
    seg = args.config == 2
    movable = synth.MOVABLE_CLASSES if seg else ()

    # synthetic data: each rank owns its own sequences (config 3 sharding: seq_id % world == rank) and generates the first of
    # them: POOL consecutive frame pairs.  Step s takes pairs [s*B, s*B + B) mod POOL -- a sliding window over the sequence.
    seqs = my_sequences(64, rank, world)
    pool = synth.make_sequence((2000 if seg else 1000) + seqs[0], POOL, N)
    p1 = np.stack([it["pos1"] for it in pool])
    p2 = np.stack([it["pos2"] for it in pool])
    if seg:
        h_sem = np.stack([it["sem"] for it in pool]).astype(np.int32)
        h_inst = np.stack([it["inst"] for it in pool]).astype(np.int32)
        n_inst = int(h_inst.max()) + 1

    def batch_ids(step):
        return [(step * B + i) % POOL for i in range(B)]

    n_distinct = 1
    while n_distinct < K + Wm and (n_distinct * B) % POOL != 0:   # the window sequence repeats after POOL / gcd(B, POOL) steps
        n_distinct += 1

    sd = random_init_state_dict(0)
    net = TFlow()
    net.load_state_dict(sd, strict=True)
    NS = max(1, args.streams)
    fe = SceneFlowFrontEnd(net, device=dev, tau=0.10, movable=movable, n_slots=NS, use_graph=args.graph, masker=args.masker)   # also prepares the weight images
    d1, d2 = torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)
    dev_batches = []
    for s in range(n_distinct):
        ids = batch_ids(s)
        item = [d1[ids].contiguous(), d2[ids].contiguous()]
        if seg:
            item += [torch.from_numpy(h_sem[ids]).to(dev), torch.from_numpy(h_inst[ids]).to(dev)]
        dev_batches.append(item)

    # Independent batches are pipelined over NS CUDA streams (frame pairs carry no cross-batch state): one batch's
    # latency-bound kernels (FPS: 128 CTAs x 1.7 ms) and persistent-kernel tails run under the other batch's dense kernels.
    prio = os.environ.get("SSF_STREAM_PRIO")   # experiment: distinct stream priorities (e.g. "-2,-1,0")
    if prio:
        pl = [int(v) for v in prio.split(",")]
        streams = [torch.cuda.Stream(device=dev, priority=pl[i % len(pl)]) for i in range(NS)]
    else:
        streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
    gatherer = ResultGatherer(dev) if world > 1 else None

    def mask_pose(item, flow, masker):
        x1 = item[0]
        if masker == "gmm":
            return F_.frontend(x1, flow, mode=0, in_mask=F_.gmm_mask(x1, flow))
        if seg:
            return F_.frontend(x1, flow, mode=1, sem=item[2], movable=movable, inst=item[3], n_inst=n_inst, tau=0.10)
        return F_.frontend(x1, flow, mode=1, tau=0.10)

    def device_step(s, keep, stream=None, masker=None):
        item = dev_batches[s % len(dev_batches)]
        with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(dev)):
            flows, _ = net.forward_pm(item[0], item[1])
            mask, odom = mask_pose(item, flows[0], masker or args.masker)
        keep.append((mask, odom))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_all(keep, main_st):
        """The job's only communication: every rank's poses + masks of all its batches to every rank, ONE packed
        all_gather_into_tensor on the gather stream once the batches are done (once per sequence, off the critical
        streams).  Per-batch gathers were measured and rejected: NCCL's CTAs busy-wait for the peer while holding SM resources
        the persistent tensor kernels need (2 GPUs: 97.4 % instead of 99 % weak-scaling efficiency)."""
        for st in streams:
            main_st.wait_stream(st)
        odom = torch.stack([o for _, o in keep])
        mask = torch.stack([m for m, _ in keep])
        gatherer.push(odom, mask, main_st)
        return gatherer.finish(main_st)

    def timed_device_leg(masker):
        keep = []
        for s in range(Wm):
            device_step(s, keep, streams[s % NS], masker)
        if gatherer is not None:
            gather_all(keep, torch.cuda.current_stream(dev))   # warm-up includes the gather (NCCL channels, buffers)
        sync_all()
        if gatherer is not None:
            gatherer.gather_ms()
        keep.clear()
        l0 = nat.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        main_st = torch.cuda.current_stream(dev)
        ev0.record(main_st)
        for st in streams:
            st.wait_event(ev0)
        for s in range(K):
            device_step(Wm + s, keep, streams[s % NS], masker)
        for st in streams:
            main_st.wait_stream(st)
        if gatherer is not None:
            gather_all(keep, main_st)   # the job is done when every rank holds every pose and mask
        ev1.record(main_st)
        sync_all()
        launches = nat.launch_count() - l0
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        gms = torch.tensor([gatherer.gather_ms() if gatherer is not None else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        keep.clear()
        return float(ms.item()), float(gms.item()), launches

    # ---- device-resident throughput
    clocks = ClockSampler(local)
    clocks.start()
    ms, gather_ms, launches = timed_device_leg(args.masker)
    clk = clocks.stop()
    main_st = torch.cuda.current_stream(dev)

    # ---- end to end from host buffers (H2D + kernels + D2H every step), pipelined over the NS slots
    host_batches = []
    for s in range(n_distinct):
        ids = batch_ids(s)
        hb = dict(pos1=torch.from_numpy(p1[ids]), pos2=torch.from_numpy(p2[ids]))
        if seg:
            hb.update(sem=torch.from_numpy(h_sem[ids]), inst=torch.from_numpy(h_inst[ids]), n_inst=n_inst)
        host_batches.append(hb)

    def submit(s, slot):
        return fe.submit(slot=slot, **host_batches[s % len(host_batches)])

    for s in range(Wm):
        submit(s, s % NS)
    for slot in range(NS):
        if fe._pending[slot] is not None:
            fe._pending[slot].result()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_st)
    for st in fe._streams:
        st.wait_event(e0)
    pend = [None] * NS
    for s in range(K):
        slot = s % NS
        if pend[slot] is not None:
            out = pend[slot].result()
        pend[slot] = submit(Wm + s, slot)
    for p_ in pend:
        if p_ is not None:
            out = p_.result()
    for st in fe._streams:
        main_st.wait_stream(st)
    e1.record(main_st)
    sync_all()
    ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    ems = float(ems.item())

    other = "gmm" if args.masker == "residual" else "residual"
    other_ms = latency_ms = None
    shares, point_ops, maskers = {}, None, None
    if not args.no_extras:
        # ---- the same device-resident leg with the other masker (so both arms can be compared under either)
        if not seg:
            other_ms, _, _ = timed_device_leg(other)

        # ---- single-pair latency (the reference's operating point: one frame pair per 100 ms tick), host buffers in and out
        fe1 = SceneFlowFrontEnd(net, device=dev, tau=0.10, movable=movable, n_slots=1, use_graph=args.graph, masker=args.masker)
        one = {k: (v[:1] if hasattr(v, "shape") else v) for k, v in host_batches[0].items()}
        for _ in range(5):
            fe1.process(**one)
        torch.cuda.synchronize()
        t_lat = time.perf_counter()
        for _ in range(20):
            fe1.process(**one)
        latency_ms = (time.perf_counter() - t_lat) / 20 * 1e3
        del fe1

    # ---- per-kernel shares (CUDA events around every launch of ours, same workload, after the timed region; on one of the
    # pipeline streams so that the caching allocator's warm pool is used: a cold pool would put cudaMalloc inside the events)
    keep = []
    with torch.cuda.stream(streams[0]):
        prof.enable()
        for s in range(min(K, 3)):
            device_step(Wm + s, keep)
        torch.cuda.synchronize()
        shares = prof.summary()
        prof.disable()
    keep.clear()

    if not args.no_extras and rank == 0 and not seg:
        point_ops = point_op_rooflines(B, N, dev)
        # ---- the two maskers stand-alone on the step's batch (ground-truth flow + 2 cm noise, so that there are movers to find)
        gt = np.stack([pool[i]["gt"] for i in batch_ids(0)])
        gflow = torch.from_numpy((gt + np.random.default_rng(0).normal(0, 0.02, gt.shape)).astype(np.float32)).to(dev)
        x1 = dev_batches[0][0]
        maskers = {}
        for name, fn in (("residual", lambda: F_.frontend(x1, gflow, mode=1, tau=0.10)),
                         ("gmm", lambda: F_.frontend(x1, gflow, mode=0, in_mask=F_.gmm_mask(x1, gflow)))):
            t = _timed(fn, reps=5, warm=2)
            maskers[name] = {"ms_per_batch": t * 1e3, "clouds": B, "bytes_per_cloud": N * 25 + 56}
        _, info = F_.gmm_mask(x1, gflow, want_info=True)
        maskers["gmm"]["em_iterations_mean"] = float(info[:, 0].mean().item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if args.shares_out:
        with open(args.shares_out, "w") as f:
            json.dump({k: {"ms_per_step": v["ms"] / min(K, 3), "calls_per_step": v["calls"] / min(K, 3), "share": v["share"]}
                       for k, v in sorted(shares.items(), key=lambda kv: -kv[1]["ms"])}, f, indent=1)

    pk = peaks()
    top = max(shares.items(), key=lambda kv: kv[1]["ms"])[0] if shares else None
    roof = None
    if top is not None:
        t = shares[top]
        import re
        dims = dict((k, int(v)) for k, v in re.findall(r"(\w+)=(\d+)", top))
        avg_s = t["ms"] / t["calls"] * 1e-3
        if top.startswith("cost_volume"):
            ach = B * cost_volume_flops(dims["N1"], dims["m"]) / avg_s / 1e12
            roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s",
                    "frac": ach / pk["tensor"], "traffic": None, "peak_source": pk["source"],
                    "note": ("tcgen05 kind::tf32 with the fp32-faithful 3xTF32 split (3 MMA passes per algorithmic MAC) for m=64, "
                             "attention inside the same kernel; achieved = algorithmic 2*MAC of one launch over "
                             "B clouds / its mean CUDA-event duration; peak = measured dense bf16 cuBLAS throughput")}
        elif top.startswith("dense_tc"):
            ach = 2.0 * dims["rows"] * dims["K"] * dims["N"] / avg_s / 1e12
            roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s",
                    "frac": ach / pk["tensor"], "traffic": None, "peak_source": pk["source"],
                    "note": "tcgen05 kind::tf32, 3xTF32 split; achieved = 2*rows*K*N of one launch / its mean CUDA-event duration"}
        elif top.startswith("knn"):
            byts = B * (12.0 * (dims["Nq"] + dims["Nr"]) + 4.0 * dims["Nq"] * dims["k"])
            roof = {"kernel": top, "bound": "hbm", "achieved": byts / avg_s / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": byts / avg_s / 1e9 / pk["hbm"], "traffic": None, "peak_source": pk["source"],
                    "note": "the exact block-pruned kNN is instruction-issue bound (shuffle / compare chains), not HBM bound; bytes are compulsory traffic"}
        else:
            roof = {"kernel": top, "bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None,
                    "traffic": None, "peak_source": pk["source"]}
        tr = traffic_table().get(top)   # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same batch only)
        if tr is not None and tr.get("batch"):
            roof["traffic"] = (tr["dram_read_bytes"] + tr["dram_write_bytes"]) * B / tr["batch"]
            roof["traffic_source"] = tr["capture"] + ("" if tr["batch"] == B else "; captured at %d clouds per launch, scaled to %d (the kernel's traffic is linear in the number of clouds)" % (tr["batch"], B))

    value = world * B * K / (ms * 1e-3)
    e2e = world * B * K / (ems * 1e-3)
    line = {"metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic CARLA-shaped clouds (ssf_slam_b200/synth.py), random-init TFlow weights (seed 0)",
            "config": {"workload": workload_name(args), "masker": args.masker,
                       "pairs_per_step_per_gpu": B, "distinct_pairs": POOL, "distinct_batches": n_distinct,
                       "batching": "step s = frame pairs [s*B, s*B+B) mod %d of the rank's 200-frame sequence (sliding window)" % POOL,
                       "sharding": "sequence id %% world (config 3), no collective on the hot path; the job's poses+masks are "
                                   "all-gathered once, by one packed all_gather_into_tensor on a side stream, inside the timed window",
                       "pipelining": "independent batches alternate over %d CUDA streams (per-stream pinned staging in the e2e leg)" % NS,
                       "l2": "per-step working set (~%.1f GB of intermediates) exceeds the 126 MB L2" % (B * 0.12 * N / 8192.0)},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": fe.h2d_bytes(B, N, seg=seg), "d2h_bytes_per_step": fe.d2h_bytes(B, N),
                    "ms_per_step": ems / K},
            "gather": {"ms": gather_ms, "collectives": 1 if world > 1 else 0, "bytes_per_rank": K * B * (N + 56),
                       "note": "max over ranks of the CUDA-event duration of the one packed all_gather_into_tensor (NCCL, own stream, "
                               "warmed in the warm-up); it is inside the timed window"},
            "latency": None if latency_ms is None else {"pairs": 1, "ms_per_pair": latency_ms, "note": "SceneFlowFrontEnd.process on one "
                        "host-resident frame pair (H2D + %s + D2H), mean of 20" % ("CUDA-graph replay" if args.graph else "eager launches")},
            "other_masker": None if other_ms is None else {"masker": other, "value": world * B * K / (other_ms * 1e-3), "ms_per_step": other_ms / K,
                                                           "note": "device-resident leg repeated with the other masker"},
            "roofline": roof, "point_ops": point_ops, "maskers": maskers,
            "kernel_shares": {k: round(v["share"], 4) for k, v in sorted(shares.items(), key=lambda kv: -kv[1]["ms"])[:8]}}

    if not args.no_cpu_baseline and world == 1:
        import torch as _t
        cpu_pipeline_seconds(pool, 1, sd, args.masker, seg)  # warm-up
        t = cpu_pipeline_seconds(pool, args.cpu_sample, sd, args.masker, seg, start=1)
        line["cpu_baseline"] = {"value": len(t) / float(sum(t)), "unit": UNIT, "cores": _t.get_num_threads(), "kind": "port",
                                "sample": "%d consecutive frame pairs of the sequence (N=%d), one at a time: oracle TFlow port + %s"
                                          % (len(t), N, "sklearn GMM mask + slove_RT_by_SVD" if args.masker == "gmm" else
                                             "residual masker spec (NumPy) incl. pose")}
        if not seg and not args.no_extras:
            t2 = cpu_pipeline_seconds(pool, 4, sd, other, seg, start=1)
            line["cpu_baseline"]["other_masker"] = {"masker": other, "value": len(t2) / float(sum(t2)), "sample": "4 frame pairs"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
