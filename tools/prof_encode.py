#!/usr/bin/env python3
"""Tiny driver for ncu: code one synthetic 4K frame with nh_encode_frame (config 3 or 5)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_configs import synth_plane  # noqa: E402
from nano_hevc_b200 import batched  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=16)
ap.add_argument("--cost", default="sad")
ap.add_argument("--wavefront", action="store_true")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--frames", type=int, default=1, help="frames of the batch (one nh_encode_frames call)")
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--width", type=int, default=3840)
args = ap.parse_args()
dev = torch.device("cuda:0")
planes = torch.stack([synth_plane(args.height, args.width, i, dev) for i in range(args.frames)])
for _ in range(args.reps):
    r = batched.encode_frames(planes, args.size, cost=args.cost, qp=27, recon_neighbours=args.wavefront)
torch.cuda.synchronize()
print("modes histogram:", torch.bincount(r.modes.reshape(-1).to(torch.int64), minlength=35).tolist())
