"""Runs warm-up forwards, then ONE forward of the scene-flow front end between cudaProfilerStart/Stop (for
`ncu --profile-from-start off`).  Usage: python scripts/prof_forward.py [--batch B] [--npoints N]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ssf_slam_b200 import functional as F_, synth
from ssf_slam_b200.model import TFlow
from ssf_slam_b200.weights import random_init_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--npoints", type=int, default=8192)
args = ap.parse_args()
pool = synth.make_sequence(1000, args.batch, args.npoints)
x1 = torch.from_numpy(np.stack([it["pos1"] for it in pool])).cuda()
x2 = torch.from_numpy(np.stack([it["pos2"] for it in pool])).cuda()
net = TFlow()
net.load_state_dict(random_init_state_dict(0), strict=True)
for _ in range(2):
    flows, _ = net.forward_pm(x1, x2)
    F_.frontend(x1, flows[0], mode=1, tau=0.10)
torch.cuda.synchronize()
torch.cuda.profiler.start()
flows, _ = net.forward_pm(x1, x2)
F_.frontend(x1, flows[0], mode=1, tau=0.10)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(flows[0].abs().max()))
