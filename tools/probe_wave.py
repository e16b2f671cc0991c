#!/usr/bin/env python3
"""Where does the wavefront coder's time go?  Times nh_encode_frames(recon_neighbours=1) on planes whose shape
isolates one term of  t = bw * T + bh * (2 T + L):
   one block row   (H = N, W wide)    -> T, the dependent time of one block (no inter-row wait)
   one block column (W = N.., H tall) -> 2 T + L per row
   a 4K frame, and F frames in one call (throughput).
Prints one JSON line per case."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_configs import synth_plane, time_ms  # noqa: E402
from nano_hevc_b200 import batched  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="4,8,16,32")
ap.add_argument("--cost", default="sad")
ap.add_argument("--frames", default="1,8,32")
args = ap.parse_args()
dev = torch.device("cuda:0")
big = synth_plane(4320, 7680, 0, dev)
for n in [int(v) for v in args.sizes.split(",")]:
    cases = [("one_row", n, 7680), ("two_rows", 2 * n, 7680), ("narrow", 4320 // n * n, 4 * n), ("4k", 2160, 3840)]
    for name, H, W in cases:
        p = big[:H, :W].contiguous().unsqueeze(0)
        res = batched.encode_frames(p, n, cost=args.cost, qp=27, recon_neighbours=True)
        scratch = torch.empty((1 << 26,), dtype=torch.uint8, device=dev)
        ms = time_ms(lambda: batched.encode_frames(p, n, cost=args.cost, qp=27, recon_neighbours=True, out=res, scratch=scratch), 3, warmup=1)
        bw, bh = W // n, H // n
        print(json.dumps({"N": n, "case": name, "H": H, "W": W, "bw": bw, "bh": bh, "ms": ms,
                          "us_per_block_col": ms * 1e3 / bw, "us_per_block_row": ms * 1e3 / bh,
                          "us_per_step": ms * 1e3 / (bw + 2 * bh)}), flush=True)
    for F in [int(v) for v in args.frames.split(",")]:
        p = torch.stack([synth_plane(2160, 3840, i, dev) for i in range(F)])
        res = batched.encode_frames(p, n, cost=args.cost, qp=27, recon_neighbours=True)
        scratch = torch.empty((1 << 28,), dtype=torch.uint8, device=dev)
        ms = time_ms(lambda: batched.encode_frames(p, n, cost=args.cost, qp=27, recon_neighbours=True, out=res, scratch=scratch), 2, warmup=1)
        px = F * (2160 // n) * (3840 // n) * n * n
        print(json.dumps({"N": n, "case": f"4k_x{F}", "ms": ms, "Gpix_s": px / ms / 1e6}), flush=True)
        del p, res, scratch
