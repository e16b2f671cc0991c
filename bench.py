#!/usr/bin/env python3
"""bench.py -- Mpix/s of the batched intra + transform + quant pipeline (BASELINE.json metric).

Workload (config.workload = "cfg2"): BASELINE config 2 -- synthetic 8-bit 1080p luma frames
tiled into 8x8 blocks, DC and planar prediction from given reference samples, DCT, quantise /
dequantise, inverse, reconstruct at QP 22/27/32/37.  One *step* = 8 passes (2 modes x 4 QPs) of
the fused kernel over a batch of F frames per GPU; 1 px = one luma sample pushed once through
predict -> residual -> forward -> quant -> dequant -> inverse -> recon.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

  value     device-resident throughput (inputs already in HBM), CUDA events, max over ranks.
  e2e       same metric through the host-buffer C-ABI call (nh_host_pipeline_dcplanar): pinned
            host inputs -> H2D -> kernel -> D2H of all four outputs, every pass, inside the timing.
  roofline  fused 8x8 kernel: algorithmic bytes (14.5625 B/px) / mean launch time
            against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle port (oracle/nh_oracle.c) on the host cores, bounded sample.

--impl reference times the reference's CPU implementation of the same path (the oracle port of
its numpy functions; the reference itself is pure Python and does not travel to the GPU box) on
all host threads and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, N = 1080, 1920, 8
BLOCKS_PER_FRAME = (H // N) * (W // N)      # 32,400
PX_PER_FRAME = BLOCKS_PER_FRAME * N * N     # 2,073,600
QPS = (22, 27, 32, 37)
MODES = (1, 0)                              # DC, planar
PASSES = len(QPS) * len(MODES)
# orig 2 + pred 2 + coeff 4 + levels 4 + recon 2 B/px + refs (2N+2)*2 B / N^2 px
BYTES_PER_PX = 14.0 + (2 * N + 2) * 2 / (N * N)   # 14.5625
METRIC = "Mpix/s of batched intra+transform+quant pipeline"
UNIT = "Mpix/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner
# on fd 1 when NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the whole run
# and the JSON line goes to a duplicate of the original stdout.
_JSON_FD = None


def claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def synth_frames(n_frames, seed):
    """SURVEY.md 8d (iii) 'smooth' synthetic luma: separable ramp + seeded low-amplitude noise."""
    rng = np.random.default_rng(4321 + seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = (40 + (150 * xx) // (W - 1) + (60 * yy) // (H - 1)).astype(np.int16)
    out = np.empty((n_frames, H, W), np.int16)
    for f in range(n_frames):
        out[f] = np.clip(base + rng.integers(-12, 13, (H, W), dtype=np.int16), 0, 255)
    return out


def frame_to_cfg2_inputs(frames):
    """(F,H,W) int16 -> block-major orig and the given references with the CLI convention
    (N samples per side, 128 at frame edges, top_right = top[-1], bottom_left = left[-1])."""
    F = frames.shape[0]
    bh, bw = H // N, W // N
    orig = np.ascontiguousarray(frames.reshape(F, bh, N, bw, N).transpose(0, 1, 3, 2, 4)).reshape(F * bh * bw, N, N)
    top = np.full((F, bh, bw, N), 128, np.int16)
    left = np.full((F, bh, bw, N), 128, np.int16)
    above = frames[:, N - 1:H - 1:N, :]                      # row y-1 of every block row >= 1
    top[:, 1:] = above.reshape(F, bh - 1, bw, N)
    leftcol = frames[:, :, N - 1:W - 1:N]                    # column x-1 of every block col >= 1
    left[:, :, 1:] = leftcol.reshape(F, bh, N, bw - 1).transpose(0, 1, 3, 2)
    top = top.reshape(-1, N)
    left = left.reshape(-1, N)
    return orig, top, left, np.ascontiguousarray(top[:, -1]), np.ascontiguousarray(left[:, -1])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------ CPU legs
def cpu_port_rate(n_frames, threads):
    """Oracle port (oracle/nh_oracle.c) over `n_frames` frames x 8 passes on `threads` host threads."""
    import oracle as O
    frames = synth_frames(n_frames, 7)
    orig, top, left, tr, bl = frame_to_cfg2_inputs(frames)
    O.pipeline_dcplanar_batch(orig[:256], top[:256], left[:256], tr[:256], bl[:256], 1, 22)  # build + warm
    t0 = time.perf_counter()
    for mode in MODES:
        for qp in QPS:
            O.pipeline_dcplanar_batch(orig, top, left, tr, bl, mode, qp, threads=threads)
    dt = time.perf_counter() - t0
    return n_frames * PX_PER_FRAME * PASSES / dt / 1e6, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle as O
    threads = O.n_host_threads()
    sample_frames = max(2, min(128, 2 * threads))
    for _ in range(args.warmup):
        cpu_port_rate(1, threads)
    t_total, px_total = 0.0, 0
    for _ in range(args.steps):
        rate, dt = cpu_port_rate(sample_frames, threads)
        t_total += dt
        px_total += sample_frames * PX_PER_FRAME * PASSES
    value = px_total / t_total / 1e6
    sample = f"{sample_frames} frames x {PASSES} passes per step (of the cfg2 batch), {threads} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": "cfg2", "frame": f"{W}x{H}", "block": N, "modes": ["dc", "planar"], "qps": list(QPS),
                   "note": "reference's CPU implementation of the path = oracle port of its numpy functions "
                           "(the pure-Python reference does not travel to the GPU box)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# ------------------------------------------------------------------- GPU leg
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from nano_hevc_b200 import _lib, batched

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.lib().nh_device_ok() == 1, _lib.last_error()
    _lib.check(_lib.lib().nh_set_fused_impl(args.fused_impl))

    F = args.frames
    B = F * BLOCKS_PER_FRAME
    log(f"[rank {rank}] building {F} synthetic 1080p frames ({B} blocks of {N}x{N})")
    base = synth_frames(min(F, 8), rank)
    h_in = frame_to_cfg2_inputs(base)
    reps = (F + base.shape[0] - 1) // base.shape[0]
    d_in = [torch.from_numpy(a).to(dev).repeat(*([reps] + [1] * (a.ndim - 1)))[:B].contiguous() for a in h_in]
    out = batched.PipelineResult(torch.empty((B, N, N), dtype=torch.int16, device=dev),
                                 torch.empty((B, N, N), dtype=torch.int32, device=dev),
                                 torch.empty((B, N, N), dtype=torch.int32, device=dev),
                                 torch.empty((B, N, N), dtype=torch.int16, device=dev))

    def step():
        for mode in MODES:
            for qp in QPS:
                batched.fused_block_pipeline(*d_in, mode, qp, out=out)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    # statistics of the last pass: the only cross-GPU exchange, after the timed region
    stats = torch.stack([batched.count_nonzero_batched(out.levels),
                         batched.sse_sad(d_in[0], out.recon)[0]]).to(torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        stats = torch.stack(gathered).sum(0)
    total_ms = float(ms.item())
    launches = args.steps * PASSES
    px_step_all = world * F * PX_PER_FRAME * PASSES
    value = px_step_all * args.steps / (total_ms / 1e3) / 1e6

    # roofline of the dominant kernel (the only kernel in the timed region): per-launch figures
    px_launch = F * PX_PER_FRAME
    alg_bytes = BYTES_PER_PX * px_launch
    launch_s = total_ms / 1e3 / launches
    achieved = alg_bytes / launch_s / 1e9
    peak, peak_src = measured_peak()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_fused8.json")
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath))
            traffic = t["dram_bytes_per_px"] * px_launch
        except Exception:
            traffic = None

    # ---- e2e: host buffers through the C ABI, H2D + kernel + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        os.environ.setdefault("NH_HOST_THREADS", str(max(2, min(8, (os.cpu_count() or 8) // max(world, 1)))))
        Fe = min(F, args.e2e_frames)
        Be = Fe * BLOCKS_PER_FRAME
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        he = frame_to_cfg2_inputs(synth_frames(Fe, 100 + rank))
        h_in_p = [pin(a) for a in he]
        h_out = [torch.empty((Be, N, N), dtype=dt).pin_memory() for dt in (torch.int16, torch.int32, torch.int32, torch.int16)]
        chunk = int(os.environ.get("NH_E2E_CHUNK", 128 * 1024))  # measured best of 4K..128K (profiles/r1_notes.md)
        L = _lib.lib()
        sbytes = int(L.nh_host_pipeline_scratch_bytes(N, chunk))
        scratch = torch.empty((sbytes,), dtype=torch.uint8, device=dev)

        def e2e_step():
            for mode in MODES:
                for qp in QPS:
                    _lib.check(L.nh_host_pipeline_dcplanar(
                        *[t.data_ptr() for t in h_in_p], None, mode, Be, N, qp, 1, 0, 8,
                        *[t.data_ptr() for t in h_out], scratch.data_ptr(), sbytes, chunk))

        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ksteps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        delivered = PASSES * sum(t.numel() * t.element_size() for t in h_out)
        # bytes that really crossed PCIe, counted by the library from its cudaMemcpyAsync calls (the
        # output wire format is compact and data dependent); last call x PASSES calls per step
        import ctypes
        b_up, b_down = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(L.nh_host_pipeline_last_transfer(ctypes.byref(b_up), ctypes.byref(b_down)))
        h2d, d2h = PASSES * b_up.value, PASSES * b_down.value
        e2e = {"value": world * Fe * PX_PER_FRAME * PASSES * ksteps / float(dt.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_bytes_delivered_per_step": delivered,
               "frames_per_pass": Fe, "steps": ksteps,
               "api": "nh_host_pipeline_dcplanar (C ABI, pinned host buffers, 3-stream chunked overlap; compact wire format: int8 coefficients + int16 exception segments, all-zero level segments elided, widened / zero-filled on host threads; d2h bytes are those of the last pass)",
               "host_threads": int(os.environ["NH_HOST_THREADS"])}
        # spot-check the e2e outputs against the device-resident path
        chk = batched.fused_block_pipeline(*[t.to(dev) for t in h_in_p], MODES[-1], QPS[-1])
        assert torch.equal(chk.levels.cpu(), h_out[2]) and torch.equal(chk.recon.cpu(), h_out[3]), "e2e mismatch"

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            import oracle as O
            thr = O.n_host_threads()
            frames_cpu = max(2, min(128, 4 * thr))  # ~15-30 core-seconds of CPU work
            rate, dt = cpu_port_rate(frames_cpu, thr)
            cpu = {"value": rate, "unit": UNIT, "cores": thr, "kind": "port",
                   "sample": f"{frames_cpu} frames x {PASSES} passes of the same workload, {dt:.1f} s wall"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "cfg2", "frame": f"{W}x{H}", "block": N, "frames_per_gpu": F,
                       "blocks_per_gpu": B, "modes": ["dc", "planar"], "qps": list(QPS), "passes_per_step": PASSES,
                       "l2": f"inputs+outputs {alg_bytes / 1e9:.2f} GB per launch >> 126 MB L2 (no flush needed)",
                       "parallelism": f"frames sharded over {world} GPU(s), no collective on the hot path"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": {1: "fused_unit_kernel<8>", 2: "fused_unit_kernel_v2<8>", 3: "fused_unit_kernel_v3<8>", 4: "fused_mma8_kernel"}[args.fused_impl], "bytes_per_px": BYTES_PER_PX,
                         "px_per_launch": px_launch, "launch_ms": launch_s * 1e3, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "note": "peak is the driver's b.copy_(a) probe (1 read : 1 write); this write-dominated "
                                 "stream can run marginally faster than that probe, so frac may exceed 1"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "stats": {"nonzero_levels_last_pass": int(stats[0].item()), "sse_last_pass": int(stats[1].item())},
        }
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=128, help="1080p frames per GPU per pass")
    ap.add_argument("--e2e-frames", type=int, default=16)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fused-impl", type=int, default=4, choices=[1, 2, 3, 4],
                    help="kernel generation of the fused 4x4/8x8 kernel (1 = first generation, for A/B runs)")
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
