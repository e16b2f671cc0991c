#!/usr/bin/env python3
"""Tiny driver for ncu: a few launches of the fused DC/planar pipeline at one block size."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_hevc_b200 import _lib, batched  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=32)
ap.add_argument("--mpix", type=int, default=64)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--impl", type=int, default=4)
args = ap.parse_args()
dev = torch.device("cuda:0")
n = args.size
B = (args.mpix << 20) // (n * n)
g = torch.Generator(device=dev).manual_seed(1)
ri = lambda *s: torch.randint(0, 256, s, generator=g, device=dev, dtype=torch.int16)
orig, top, left, tr, bl = ri(B, n, n), ri(B, n), ri(B, n), ri(B), ri(B)
_lib.check(_lib.lib().nh_set_fused_impl(args.impl))
out = batched._outputs(("pred", "coeff", "levels", "recon"), B, n, dev)
for _ in range(args.reps):
    batched.fused_block_pipeline(orig, top, left, tr, bl, args.mode, 27, use_dst=(n == 4), out=out)
torch.cuda.synchronize()
print("ok", int(out.levels.ne(0).sum()))
