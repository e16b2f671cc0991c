#!/usr/bin/env python3
"""Tiny driver for ncu: a few launches of the single-stage forward / inverse transform at one size."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_hevc_b200 import batched  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=32)
ap.add_argument("--blocks", type=int, default=1 << 18)
ap.add_argument("--inverse", action="store_true")
ap.add_argument("--reps", type=int, default=4)
args = ap.parse_args()
dev = torch.device("cuda:0")
n = args.size
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randint(-255, 256, (args.blocks, n, n), generator=g, device=dev, dtype=torch.int32)
if not args.inverse:
    x = x.to(torch.int16)
for _ in range(args.reps):
    y = batched.inverse_transform_batched(x) if args.inverse else batched.forward_transform_batched(x)
torch.cuda.synchronize()
print("ok", int(y.abs().max()))
