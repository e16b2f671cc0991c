#!/usr/bin/env python3
"""bench_configs.py -- secondary measurements (NOT the driver's bench line; see bench.py).

Times every other kernel of the hot path on one B200 with CUDA events and prints one JSON object per
line: BASELINE configs 3 (4K 35-mode search), 4 (2^20-block transform / quant microbench), 5 (4K
wavefront coder) and the fused pipeline at every block size.  Each line carries the algorithmic bytes
 and the fraction of the measured HBM copy bandwidth, so the round notes can say which
kernels are HBM-bound and which are integer- or latency-bound.

    python bench_configs.py [--which fused,xform,search,wavefront] [--reps 5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nano_hevc_b200 import _lib, batched  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_ms(fn, reps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def emit(name, px, bytes_per_px, ms, **extra):
    gbs = px * bytes_per_px / (ms / 1e3) / 1e9
    print(json.dumps({"kernel": name, "Mpix_s": px / (ms / 1e3) / 1e6, "ms": ms, "bytes_per_px": bytes_per_px,
                      "GBs": gbs, "frac_hbm": gbs / peak(), **extra}), flush=True)


def synth_plane(H, W, seed, dev):
    g = torch.Generator(device=dev).manual_seed(4321 + seed)
    yy = torch.arange(H, device=dev).view(H, 1)
    xx = torch.arange(W, device=dev).view(1, W)
    base = 40 + (150 * xx) // (W - 1) + (60 * yy) // (H - 1)
    noise = torch.randint(-12, 13, (H, W), generator=g, device=dev)
    return (base + noise).clamp(0, 255).to(torch.int16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="fused,xform,search,wavefront")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    which = set(args.which.split(","))
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    assert _lib.lib().nh_device_ok() == 1
    g = torch.Generator(device=dev).manual_seed(99)

    if "fused" in which:
        for n in (4, 8, 16, 32):
            B = (1 << 28) // (n * n)  # 268 Mpix per launch
            orig = torch.randint(0, 256, (B, n, n), generator=g, device=dev, dtype=torch.int16)
            top = torch.randint(0, 256, (B, n), generator=g, device=dev, dtype=torch.int16)
            left = torch.randint(0, 256, (B, n), generator=g, device=dev, dtype=torch.int16)
            tr = torch.randint(0, 256, (B,), generator=g, device=dev, dtype=torch.int16)
            bl = torch.randint(0, 256, (B,), generator=g, device=dev, dtype=torch.int16)
            out = batched._outputs(("pred", "coeff", "levels", "recon"), B, n, dev)
            for mode in (1, 0):
                ms = time_ms(lambda: batched.fused_block_pipeline(orig, top, left, tr, bl, mode, 27,
                                                                  use_dst=(n == 4), out=out), args.reps)
                emit(f"fused_dcplanar N={n} mode={'dc' if mode else 'planar'}", B * n * n,
                     14 + (2 * n + 2) * 2 / (n * n), ms)
            del orig, top, left, tr, bl, out

    if "xform" in which:
        for n, dst in ((4, False), (4, True), (8, False), (16, False), (32, False)):
            B = 1 << 20
            x = torch.randint(-255, 256, (B, n, n), generator=g, device=dev, dtype=torch.int16)
            px = B * n * n
            tag = f"N={n}{' dst' if dst else ''}"
            ms = time_ms(lambda: batched.forward_transform_batched(x, dst), args.reps)
            emit(f"forward_transform {tag}", px, 6, ms, note="includes torch.empty of the output")
            c = batched.forward_transform_batched(x, dst)
            ms = time_ms(lambda: batched.inverse_transform_batched(c, dst), args.reps)
            emit(f"inverse_transform {tag}", px, 8, ms)
            ms = time_ms(lambda: batched.quantize_batched(c, 27, n), args.reps)
            emit(f"quantize {tag}", px, 8, ms)
            lv = batched.quantize_batched(c, 27, n)
            ms = time_ms(lambda: batched.dequantize_batched(lv, 27), args.reps)
            emit(f"dequantize {tag}", px, 8, ms)
            del x, c, lv

    H, W = 2160, 3840
    if "search" in which or "wavefront" in which:
        plane = synth_plane(H, W, 0, dev)
    if "search" in which:
        for n in (4, 8, 16, 32):
            for cost in ("sad", "satd"):
                px = (H // n) * (W // n) * n * n
                ms = time_ms(lambda: batched.encode_frame(plane, n, cost=cost, qp=27), max(2, args.reps // 2), warmup=1)
                emit(f"encode_frame search N={n} {cost} (cfg3, one 4K frame)", px, 2 + 12 + 5 / (n * n), ms)
    if "search" in which:
        # SURVEY 8d asks for a batch of F >= 32 frames: 32 frames stacked into one tall plane (source
        # neighbours, so only the first block row of each frame sees different references) show the
        # steady-state rate without the per-launch tail of a single 4K frame
        F = 32
        tall = torch.cat([synth_plane(H, W, i, dev) for i in range(F)], dim=0)
        for n in (4, 8, 16, 32):
            px = (F * H // n) * (W // n) * n * n
            for cost in ("sad", "satd"):
                ms = time_ms(lambda: batched.encode_frame(tall, n, cost=cost, qp=27), 2, warmup=1)
                emit(f"encode_frame search N={n} {cost} (cfg3, {F} 4K frames in one launch)", px, 2 + 12 + 5 / (n * n), ms)
        del tall
    if "wavefront" in which:
        for n in (4, 8, 16, 32):
            px = (H // n) * (W // n) * n * n
            ms = time_ms(lambda: batched.encode_frame(plane, n, cost="sad", qp=27, recon_neighbours=True), 2, warmup=1)
            emit(f"encode_frame wavefront N={n} sad (cfg5, one 4K frame)", px, 2 + 12 + 5 / (n * n), ms)


if __name__ == "__main__":
    main()
