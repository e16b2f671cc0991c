#!/usr/bin/env python3
"""A/B timing of the config-3 search kernels: for each block size and cost kind, code a batch of synthetic 4K frames
with every requested nh_set_search_impl() setting, check that modes and costs agree with the first setting, and print
Gpix/s (search + winner kernels, CUDA events).

    python tools/time_search.py [--sizes 16,32] [--impls 2,5] [--frames 32] [--costs sad,satd]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_configs import synth_plane, time_ms  # noqa: E402
from nano_hevc_b200 import _lib, batched  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16,32")
ap.add_argument("--impls", default="2,5")
ap.add_argument("--frames", type=int, default=32)
ap.add_argument("--costs", default="sad,satd")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda:0")
planes = torch.stack([synth_plane(2160, 3840, i, dev) for i in range(args.frames)])
px = planes.numel()
for n in [int(x) for x in args.sizes.split(",")]:
    for cost in args.costs.split(","):
        ref = None
        for impl in [int(x) for x in args.impls.split(",")]:
            _lib.check(_lib.lib().nh_set_search_impl(impl))
            r = batched.encode_frames(planes, n, cost=cost, qp=27)
            torch.cuda.synchronize()
            if ref is None:
                ref = (r.modes.clone(), r.costs.clone())
                same = True
            else:
                same = bool(torch.equal(r.modes, ref[0]) and torch.equal(r.costs, ref[1]))
            ms = time_ms(lambda: batched.encode_frames(planes, n, cost=cost, qp=27), args.reps)
            print(json.dumps({"N": n, "cost": cost, "impl": impl, "frames": args.frames, "ms": round(ms, 4),
                              "Gpix_s": round(px / ms / 1e6, 2), "same_as_first": same}), flush=True)
_lib.check(_lib.lib().nh_set_search_impl(2))
