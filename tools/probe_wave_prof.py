#!/usr/bin/env python3
"""Per-phase cycles of the wavefront kernels' first block row (needs a library built with -DNH_WAVE_PROF:
   make -C nano_hevc_b200/csrc prof   ->  nano_hevc_b200/libnh_b200_prof.so,  run with NH_B200_LIB=... )."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_configs import synth_plane  # noqa: E402
from nano_hevc_b200 import batched  # noqa: E402

dev = torch.device("cuda:0")
big = synth_plane(4320, 7680, 0, dev)
names = ["poll", "refs", "barrier1", "search", "barrier2", "winner_predict", "mma_chain", "publish+loop"]
names_mw = ["refs+barrier1", "pixels+neg+vote", "search", "barrier(keys)", "winner+publish (warp 0)", "barrier(recon)",
            "recon store", "loop head"]
for n in (16, 32):   # multi-warp kernels of nh_frame.cu (NH_WAVE_WARPS picks the warps per row)
    for name, H, W in (("one_row", n, 7680), ("4k", 2160, 3840)):
        p = big[:H, :W].contiguous().unsqueeze(0)
        res = batched.encode_frames(p, n, qp=27, recon_neighbours=True)
        scratch = torch.zeros((1 << 26,), dtype=torch.uint8, device=dev)
        batched.encode_frames(p, n, qp=27, recon_neighbours=True, out=res, scratch=scratch)
        torch.cuda.synchronize()
        c = scratch[64:128].view(torch.int64).cpu().tolist()
        bw = W // n
        print(json.dumps({"N": n, "case": name, "cycles_per_block": {k: round(v / bw, 1) for k, v in zip(names_mw, c)},
                          "total": round(sum(c) / bw, 1)}), flush=True)
for n in (8, 4):  # NH_WAVE4=1 profiles the one-warp N = 4 kernel
    for name, H, W in (("one_row", n, 7680), ("4k", 2160, 3840)):
        p = big[:H, :W].contiguous().unsqueeze(0)
        res = batched.encode_frames(p, n, qp=27, recon_neighbours=True)
        scratch = torch.zeros((1 << 26,), dtype=torch.uint8, device=dev)
        batched.encode_frames(p, n, qp=27, recon_neighbours=True, out=res, scratch=scratch)
        torch.cuda.synchronize()
        c = scratch[64:128].view(torch.int64).cpu().tolist()
        bw = W // n
        print(json.dumps({"N": n, "case": name, "cycles_per_block": {k: round(v / bw, 1) for k, v in zip(names, c)},
                          "total": round(sum(c) / bw, 1)}), flush=True)
