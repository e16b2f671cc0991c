#!/usr/bin/env python3
"""Single-stage transform kernels (N = 4 / 8) over batch sizes, rotating over buffers > 4x the L2 so that every
launch streams from HBM.  Run once with NH_XF_PIPE=0 and once without to compare the two kernels."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_hevc_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
peak = 6455.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
st = torch.cuda.current_stream().cuda_stream
for n, dst in ((4, 0), (4, 1), (8, 0)):
    for lb in (20, 22, 24):
        B = 1 << lb
        per = B * n * n * 6
        rot = max(2, -(-(4 * 126 * (1 << 20)) // per) + 1)
        xs = [torch.randint(-255, 256, (B, n, n), device=dev, dtype=torch.int16) for _ in range(rot)]
        co = [torch.empty((B, n, n), dtype=torch.int32, device=dev) for _ in range(rot)]
        rs = [torch.empty((B, n, n), dtype=torch.int32, device=dev) for _ in range(rot)]

        def fwd():
            for x, c in zip(xs, co):
                L.nh_forward_transform(x.data_ptr(), 0, c.data_ptr(), B, n, dst, st)

        def inv():
            for c, r in zip(co, rs):
                L.nh_inverse_transform(c.data_ptr(), r.data_ptr(), B, n, dst, st)

        for name, fn, bpp in (("forward", fwd, 6), ("inverse", inv, 8)):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3 / rot
            gbs = B * n * n * bpp / ms / 1e6
            print(json.dumps({"kernel": f"{name} N={n}{' dst' if dst else ''}", "blocks": B, "us_per_launch": ms * 1e3,
                              "GBs": gbs, "frac_hbm": gbs / peak, "pipe": os.environ.get("NH_XF_PIPE", "1")}), flush=True)
        del xs, co, rs
