#!/usr/bin/env python3
"""PCIe and host-memory ceilings of the box, for reading the e2e figure of bench.py:
pinned H2D / D2H alone and together, and a host-side streaming widen (int16 -> int32) on N threads."""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

dev = torch.device("cuda:0")
MB = 1 << 20
n = 512 * MB
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


out = {"h2d_GBs": n / timed(h2d) / 1e9, "d2h_GBs": n / timed(d2h) / 1e9}
t = timed(both)
out["duplex_each_GBs"] = n / t / 1e9
# host widening int16 -> int32 with numpy on T threads (numpy releases the GIL in astype/copyto)
src = np.zeros(256 * MB // 2, dtype=np.int16)
dst = np.zeros(256 * MB // 2, dtype=np.int32)
for T in (1, 2, 4, 8, 16):
    parts = np.array_split(np.arange(src.size), T)
    def work(i):
        lo, hi = parts[i][0], parts[i][-1] + 1
        np.copyto(dst[lo:hi], src[lo:hi])
    def run():
        th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
        [x.start() for x in th]
        [x.join() for x in th]
    run()
    t0 = time.perf_counter()
    for _ in range(3):
        run()
    dt = (time.perf_counter() - t0) / 3
    out[f"widen_T{T}_Gelem_s"] = src.size / dt / 1e9
out["cpus"] = os.cpu_count()
print(json.dumps(out))
