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
  e2e_noise the same call on SURVEY 8d's `noise` frames (every level segment non-zero: the compact
            wire format does not apply and everything crosses PCIe as int16) -- brackets `e2e`.
  cpu_baseline  the CPU oracle port (oracle/nh_oracle.c) on the host cores, bounded sample, plus
            `reference_numpy`: the reference's OWN numpy functions (imported from
            $NANO_HEVC_REFERENCE, baseline/_ref or /root/reference when one of them is there) on a
            fixed-seed subsample of the same workload, 1 core and all cores.
  secondary BASELINE configs 3 (32 x 4K 35-mode search), 5 (4K wavefront coder, F frames per GPU in one
            nh_encode_frames call + the NCCL gather of the per-frame statistics) and 4 (2^20-block
            transform microbench on rotating buffers > 4x L2): Gpix/s, algorithmic B/px, fraction of
            the measured HBM bandwidth and the limiting pipe (from the ncu captures in profiles/).

--impl reference times the reference's CPU implementation of the same path (the oracle port of
its numpy functions, all host threads; `cpu_baseline.reference_numpy` carries the reference's own
numpy code when it is importable) and prints the same JSON line with "impl": "reference".
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


def noise_frames(n_frames, seed):
    """SURVEY.md 8d (i) 'noise': default_rng(1234 + frame_idx).integers(0, 256, (H, W))."""
    out = np.empty((n_frames, H, W), np.int16)
    for f in range(n_frames):
        out[f] = np.random.default_rng(1234 + seed + f).integers(0, 256, (H, W), dtype=np.uint8)
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


def cfg2_config(frames_per_gpu, world):
    """config of the headline workload: identical in both arms (the driver compares them)."""
    px_launch = frames_per_gpu * PX_PER_FRAME
    return {"workload": "cfg2", "frame": f"{W}x{H}", "block": N, "frames_per_gpu": frames_per_gpu,
            "blocks_per_gpu": frames_per_gpu * BLOCKS_PER_FRAME, "modes": ["dc", "planar"], "qps": list(QPS),
            "passes_per_step": PASSES,
            "l2": f"inputs+outputs {BYTES_PER_PX * px_launch / 1e9:.2f} GB per launch >> 126 MB L2 (no flush needed)",
            "parallelism": f"frames sharded over {world} GPU(s), no collective on the hot path"}


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


# The reference's own numpy functions (never copied into this repository): importable when the driver or
# the builder has placed the unmodified package under one of these roots.
def _reference_root():
    for root in (os.environ.get("NANO_HEVC_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if root and os.path.isdir(os.path.join(root, "nano_hevc")):
            return root
    return None


def _numpy_ref_worker(job):
    """One process: the cfg2 composition (README.md:55-71) on `n` 8x8 blocks with the reference's functions."""
    root, seed, n = job
    sys.path.insert(0, root)
    import nano_hevc as R
    frames = synth_frames(1, seed)
    orig, top, left, tr, bl = frame_to_cfg2_inputs(frames)
    idx = np.random.default_rng(seed).choice(orig.shape[0], n, replace=False)
    t0 = time.perf_counter()
    for b in idx:
        for mode in MODES:
            pred = (R.intra_dc_predict(top[b], left[b], N) if mode == 1 else
                    R.intra_planar_predict(top[b], left[b], int(tr[b]), int(bl[b]), N))
            res = R.residual_block(orig[b], pred)
            co = R.forward_transform(res)
            for qp in QPS:
                lv = R.quantize_block(co, qp)
                rr = R.inverse_transform(R.dequantize_block(lv, qp))
                R.clip_to_pixel_range(R.reconstruct_block(pred, rr))
    return time.perf_counter() - t0


def reference_numpy_rate(blocks_per_proc=384):
    """The reference's own numpy code on a fixed-seed subsample of cfg2:
    Mpix/s on 1 core and on all cores (one process per core).  Note the composition shares the prediction and
    the forward transform between the four QPs of a mode, as a caller of the reference would."""
    root = _reference_root()
    if root is None:
        return {"reference_numpy": None, "why": "the reference package is not present on this box "
                                                "(NANO_HEVC_REFERENCE, baseline/_ref and /root/reference probed)"}
    import multiprocessing as mp
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    px = blocks_per_proc * N * N * PASSES
    t1 = _numpy_ref_worker((root, 0, blocks_per_proc))
    with mp.get_context("spawn").Pool(cores) as pool:
        times = pool.map(_numpy_ref_worker, [(root, 1 + i, blocks_per_proc) for i in range(cores)])
    wall = max(times)   # the processes run concurrently; each times its own compute loop
    return {"reference_numpy": {"kind": "reference", "value_1core": px / t1 / 1e6, "value_allcores": cores * px / wall / 1e6,
                                "unit": UNIT, "cores": cores, "root": os.path.relpath(root, ROOT) if root.startswith(ROOT) else root,
                                "sample": f"{blocks_per_proc} blocks x {PASSES} passes per process (fixed seed), 1 process "
                                          f"({t1:.1f} s) then {cores} concurrent processes (slowest {wall:.1f} s)"}}


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
        "config": cfg2_config(args.frames, world),
        "note": "reference's CPU implementation of the path = oracle port of its numpy functions on all host "
                "threads; each step is a bounded sample of the cfg2 batch (see cpu_baseline.sample)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         **reference_numpy_rate()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# ------------------------------------------------------------------- secondary configs (GPU)
def _synth_plane_dev(torch, Hh, Ww, seed, dev):
    g = torch.Generator(device=dev).manual_seed(4321 + seed)
    yy = torch.arange(Hh, device=dev).view(Hh, 1)
    xx = torch.arange(Ww, device=dev).view(1, Ww)
    base = 40 + (150 * xx) // (Ww - 1) + (60 * yy) // (Hh - 1)
    return (base + torch.randint(-12, 13, (Hh, Ww), generator=g, device=dev)).clamp(0, 255).to(torch.int16)


def run_secondary(torch, dist, batched, dev, rank, world, peak):
    """BASELINE configs 3, 5 and 4 on every rank (weak scaling: per-GPU work fixed), CUDA events, max over
    ranks; value = work of all ranks / that time.  Limiting pipes are the ncu findings in profiles/."""
    H4, W4 = 2160, 3840

    def timed(fn, reps, warmup=1):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def entry(px_per_gpu, bpp, ms, **extra):
        gpix = world * px_per_gpu / (ms / 1e3) / 1e9
        gbs = px_per_gpu * bpp / (ms / 1e3) / 1e9          # per GPU
        return {"Gpix_s": gpix, "ms": ms, "bytes_per_px": bpp, "GBs_per_gpu": gbs, "frac_hbm": gbs / peak, **extra}

    out = {"n_gpus": world, "peak_hbm_gbs": peak,
           "note": "per-GPU work fixed (weak scaling); Gpix_s is the aggregate over all ranks, frac_hbm per GPU"}

    def e2e_frames(host_planes, n, recon_neighbours, outputs, reps=2):
        """The same coder through nh_host_encode_frames: pinned HOST frames in, pinned HOST results out, copies
        inside the timed region (wall clock around the blocking call, max over ranks)."""
        import time
        L = __import__("nano_hevc_b200")._lib.lib()
        Fh = host_planes.shape[0]
        res = batched.host_encode_frames(host_planes, n, cost="sad", qp=27, recon_neighbours=recon_neighbours,
                                         outputs=outputs, stats=True, frames_per_chunk=2, device=dev)
        nbytes = int(L.nh_host_encode_frames_scratch_bytes(2, H4, W4, n, int(recon_neighbours)))
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        call = lambda: batched.host_encode_frames(host_planes, n, cost="sad", qp=27, recon_neighbours=recon_neighbours,
                                                  outputs=outputs, stats=True, frames_per_chunk=2, device=dev,
                                                  scratch=scratch, out=res)
        call()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        import ctypes
        up, down = ctypes.c_int64(0), ctypes.c_int64(0)
        L.nh_host_encode_frames_last_transfer(ctypes.byref(up), ctypes.byref(down))
        px = Fh * (H4 // n) * (W4 // n) * n * n
        return {"Gpix_s": world * px / (ms / 1e3) / 1e9, "ms": ms, "frames_per_gpu": Fh, "outputs": list(outputs),
                "h2d_bytes_per_call": up.value, "d2h_bytes_per_call": down.value,
                "GBs_pcie_down_per_gpu": down.value / (ms / 1e3) / 1e9,
                "api": "nh_host_encode_frames (C ABI, pinned host buffers, chunks of 2 frames over 3 streams)",
                "bound": "PCIe: results cross the bus in the reference's dtypes"}

    ALL = ("modes", "costs", "pred", "coeff", "levels", "recon_planes")
    # ---- cfg3: 32 4K frames per GPU, one nh_encode_frames call (search + winner pipeline, all outputs)
    F3 = 32
    planes = torch.stack([_synth_plane_dev(torch, H4, W4, 100 * rank + i, dev) for i in range(F3)])
    cfg3 = {}
    for n in (4, 8, 16, 32):
        px = F3 * (H4 // n) * (W4 // n) * n * n
        res = batched.encode_frames(planes, n, cost="sad", qp=27, stats=False)
        for cost in ("sad", "satd"):
            ms = timed(lambda: batched.encode_frames(planes, n, cost=cost, qp=27, stats=False, out=res), 2)
            lim = ("instruction issue (search kernel: ALU pipe 55 %, tensor pipe 33 %, issue 62 %; SATD on the tensor cores, "
                   "profiles/r4_search_quad8_v3_ncu_summary.json, r4_search_quad4_v2_ncu_summary.json)" if cost == "satd"
                   else "ALU pipe (search kernel), see profiles/")
            cfg3[f"N{n}_{cost}"] = entry(px, 2 + 12 + 5 / (n * n), ms, frames_per_gpu=F3, limiter=lim)
        del res
    host8 = torch.stack([_synth_plane_dev(torch, H4, W4, 100 * rank + i, dev) for i in range(8)]).cpu().pin_memory()
    cfg3["N8_sad"]["e2e"] = e2e_frames(host8, 8, False, ALL)
    cfg3["N8_sad"]["e2e_modes_levels_recon"] = e2e_frames(host8, 8, False, ("modes", "levels", "recon_planes"))
    out["cfg3"] = cfg3
    # ---- cfg5: wavefront coder, F 4K frames per GPU in one call + one frame alone (latency); stats by NCCL
    cfg5 = {}
    for n in (4, 8, 16, 32):
        px1 = (H4 // n) * (W4 // n) * n * n
        # N = 8 / 32: one frame (latency), 8 frames (the config), 32 frames (rows in flight fill the GPU);
        # N = 4 / 16: the one-frame latency only
        for F in ((1, 8, 32) if n in (8, 32) else (1,)):
            sub = planes[:F]
            res = batched.encode_frames(sub, n, cost="sad", qp=27, recon_neighbours=True)
            scratch = torch.empty((int(1 << 26),), dtype=torch.uint8, device=dev)
            ms = timed(lambda: batched.encode_frames(sub, n, cost="sad", qp=27, recon_neighbours=True, out=res,
                                                     scratch=scratch), 2)
            e = entry(F * px1, 2 + 12 + 5 / (n * n), ms, frames_per_gpu=F,
                      limiter="dependency latency (anti-diagonal wavefront)")
            if F == 8:   # final gather of the per-frame statistics: the only collective of the config
                st = res.stats.clone()
                if world > 1:
                    g = [torch.empty_like(st) for _ in range(world)]
                    dist.all_gather(g, st)
                    st = torch.cat(g)
                sse = st[:, 0].double() / st[:, 1].double()
                e["psnr_first_frames"] = [float(v) for v in (10 * torch.log10(255.0 ** 2 / sse))[:2].tolist()]
                e["frames_total"] = int(st.shape[0])
            cfg5[f"N{n}_F{F}"] = e
            del res, scratch
    cfg5["N8_F8"]["e2e"] = e2e_frames(host8, 8, True, ALL)
    out["cfg5"] = cfg5
    del planes, host8
    # ---- cfg4: 2^20 blocks per size, forward / inverse alone, rotating over buffers > 4x the 126 MB L2
    cfg4 = {}
    Bn = 1 << 20
    g = torch.Generator(device=dev).manual_seed(99)
    for n, dst in ((4, False), (4, True), (8, False), (16, False), (32, False)):
        per_call = Bn * n * n * 6                      # int16 in + int32 out
        rot = max(2, -(-(4 * 126 * (1 << 20)) // per_call) + 1)
        xs = [torch.randint(-255, 256, (Bn, n, n), generator=g, device=dev, dtype=torch.int16) for _ in range(rot)]
        L = __import__("nano_hevc_b200")._lib.lib()
        co = [torch.empty((Bn, n, n), dtype=torch.int32, device=dev) for _ in range(rot)]
        rs = [torch.empty((Bn, n, n), dtype=torch.int32, device=dev) for _ in range(rot)]
        st = torch.cuda.current_stream().cuda_stream

        def fwd():
            for x, c in zip(xs, co):
                L.nh_forward_transform(x.data_ptr(), 0, c.data_ptr(), Bn, n, int(dst), st)

        def inv():
            for c, r in zip(co, rs):
                L.nh_inverse_transform(c.data_ptr(), r.data_ptr(), Bn, n, int(dst), st)

        tag = f"N{n}{'dst' if dst else ''}"
        cfg4[f"{tag}_forward"] = entry(rot * Bn * n * n, 6, timed(fwd, 3), rotating_buffers=rot)
        cfg4[f"{tag}_inverse"] = entry(rot * Bn * n * n, 8, timed(inv, 3), rotating_buffers=rot)
        del xs, co, rs
    out["cfg4"] = cfg4
    return out


# ------------------------------------------------------------------- GPU leg
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from nano_hevc_b200 import _lib, batched

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    # pinned buffers and the library's host threads stay on the NUMA node of this rank's GPU
    from nano_hevc_b200 import hostbind
    numa = hostbind.bind_to_gpu_numa_node(local_rank) if world > 1 else {"bound": False, "note": "single rank: not bound"}
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
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_fused8.json")
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath))
            traffic = t["dram_bytes_per_px"] * px_launch
            traffic_source = "ncu capture profiles/traffic_fused8.json (dram bytes per px of an earlier run x this run's px per launch; not measured in this run)"
        except Exception:
            traffic = None

    # ---- e2e: host buffers through the C ABI, H2D + kernel + D2H inside the timed region
    e2e = e2e_noise = None
    if not args.no_e2e:
        try:
            avail = len(os.sched_getaffinity(0))
        except AttributeError:
            avail = os.cpu_count() or 8
        # host threads of the widening pass: this rank's share of the cores the job may use
        share = numa.get("cpus_before", avail) // max(world, 1) if world > 1 else avail - 1
        # ... capped at 6: measured on the 16-vCPU box (profiles/r4_host_threads.txt), e2e 8.3 / 7.7 / 7.2 / 6.0 / 3.9 Gpix/s
        # with 4 / 6 / 8 / 10 / 15 threads -- the pass is bound by host memory, and threads beyond what saturates it
        # only delay the DMA engines' traffic (the time spent waiting for a chunk's copies grows 2.5x from 8 to 15)
        os.environ.setdefault("NH_HOST_THREADS", str(max(2, min(6, share))))
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        chunk = int(os.environ.get("NH_E2E_CHUNK", 128 * 1024))  # measured best of 4K..128K (profiles/r1_notes.md)
        L = _lib.lib()
        sbytes = int(L.nh_host_pipeline_scratch_bytes(N, chunk))
        scratch = torch.empty((sbytes,), dtype=torch.uint8, device=dev)
        import ctypes

        def e2e_leg(frames_np, ksteps, what, int16_results=False):
            Fe = frames_np.shape[0]
            Be = Fe * BLOCKS_PER_FRAME
            h_in_p = [pin(a) for a in frame_to_cfg2_inputs(frames_np)]
            wide = torch.int16 if int16_results else torch.int32
            h_out = [torch.empty((Be, N, N), dtype=dt).pin_memory() for dt in (torch.int16, wide, wide, torch.int16)]
            entry = L.nh_host_pipeline_dcplanar_i16 if int16_results else L.nh_host_pipeline_dcplanar

            def e2e_step():
                for mode in MODES:
                    for qp in QPS:
                        _lib.check(entry(
                            *[t.data_ptr() for t in h_in_p], None, mode, Be, N, qp, 1, 0, 8,
                            *[t.data_ptr() for t in h_out], scratch.data_ptr(), sbytes, chunk))

            e2e_step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
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
            b_up, b_down = ctypes.c_int64(0), ctypes.c_int64(0)
            _lib.check(L.nh_host_pipeline_last_transfer(ctypes.byref(b_up), ctypes.byref(b_down)))
            res = {"value": world * Fe * PX_PER_FRAME * PASSES * ksteps / float(dt.item()) / 1e6, "unit": UNIT,
                   "h2d_bytes_per_step": PASSES * b_up.value, "d2h_bytes_per_step": PASSES * b_down.value,
                   "host_bytes_delivered_per_step": delivered, "frames_per_pass": Fe, "steps": ksteps, "content": what,
                   "host_threads": int(os.environ["NH_HOST_THREADS"]), "host_cores_available": avail, "numa": numa}
            # spot-check the e2e outputs against the device-resident path (first 8 frames)
            nb = min(Be, 8 * BLOCKS_PER_FRAME)
            chk = batched.fused_block_pipeline(*[t[:nb].to(dev) for t in h_in_p], MODES[-1], QPS[-1])
            assert torch.equal(chk.levels.cpu(), h_out[2][:nb].to(torch.int32)) and torch.equal(chk.recon.cpu(), h_out[3][:nb]), "e2e mismatch"
            assert torch.equal(chk.coeff.cpu(), h_out[1][:nb].to(torch.int32)) and torch.equal(chk.pred.cpu(), h_out[0][:nb]), "e2e mismatch"
            return res

        Fe = min(F, args.e2e_frames)
        base_e = synth_frames(min(Fe, 16), 100 + rank)
        frames_e = np.concatenate([base_e] * (-(-Fe // base_e.shape[0])))[:Fe]
        e2e = e2e_leg(frames_e, max(1, min(args.steps, 3)), "smooth (SURVEY 8d iii), the frames of the headline config")
        e2e["api"] = ("nh_host_pipeline_dcplanar (C ABI, pinned host buffers, 3-stream chunked overlap; compact wire format: "
                      "int8 coefficients + int16 exception segments, all-zero level segments elided, widened / zero-filled "
                      "on host threads into the reference's int32 arrays; d2h bytes are those of the last pass)")
        e2e["bound"] = ("host memory + PCIe: per pixel 7.8 B cross PCIe (2.56 up, 5.26 down) and the host threads write 8 B of int32 "
                        "coefficients / levels; the DMA probe below gives the box's ceiling, profiles/r4_host_threads.txt the "
                        "thread sweep (more than ~6 widening threads only delay the DMA traffic)")
        Fn = min(Fe, 32)
        e2e_noise = e2e_leg(noise_frames(Fn, 1000 * rank), 1, "noise (SURVEY 8d i): every level segment non-zero, int16 wire format")
        # the same workload with coefficients / levels delivered as int16 (an option of the API, not the reference's
        # dtypes): DMA straight into the caller's arrays, no host pass -- what is left is PCIe
        e2e_i16 = e2e_leg(frames_e, max(1, min(args.steps, 3)), "smooth, int16 delivery of coefficients and levels", int16_results=True)
        e2e_i16["api"] = ("nh_host_pipeline_dcplanar_i16 (C ABI, pinned host buffers, 3-stream chunked overlap; pred / coeff / levels / "
                          "recon by DMA into the caller's int16 arrays: 8 B/px down, no host threads)")
        e2e_i16["bound"] = "PCIe (2.56 B/px up, 8 B/px down)"
        e2e["int16_delivery"] = e2e_i16

        # ---- the box's DMA ceiling: every rank copies pinned host <-> device in both directions AT THE SAME TIME, in
        # the up : down ratio of the int16-delivery path (1 : 3), nothing else running.  What the ranks reach together
        # is what the host's memory / IO fabric gives the e2e legs; per-rank share x bytes per pixel = their ceiling.
        def dma_probe():
            n_up, n_down = 256 << 20, 768 << 20
            h_up = torch.empty(n_up, dtype=torch.uint8).pin_memory()
            h_down = torch.empty(n_down, dtype=torch.uint8).pin_memory()
            d_up = torch.empty(n_up, dtype=torch.uint8, device=dev)
            d_down = torch.empty(n_down, dtype=torch.uint8, device=dev)
            s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

            def both():
                with torch.cuda.stream(s1):
                    d_up.copy_(h_up, non_blocking=True)
                with torch.cuda.stream(s2):
                    h_down.copy_(d_down, non_blocking=True)

            both()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                both()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            per_rank = 3 * (n_up + n_down) / float(dt.item()) / 1e9
            return {"per_rank_GBs": per_rank, "all_ranks_GBs": per_rank * world, "ranks": world,
                    "pattern": "pinned H2D 256 MB + D2H 768 MB concurrently on every rank, 3 rounds",
                    "int16_delivery_ceiling_Mpix_s": world * per_rank * 1e9 / 10.56 / 1e6,
                    "note": "int16 delivery moves 2.56 + 8 = 10.56 B/px over PCIe; its ceiling is this aggregate DMA rate / 10.56"}

        e2e["dma_ceiling"] = dma_probe()
        del scratch

    secondary = None
    if not args.no_secondary:
        secondary = run_secondary(torch, dist, batched, dev, rank, world, peak)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            import oracle as O
            thr = O.n_host_threads()
            frames_cpu = max(2, min(128, 4 * thr))  # ~15-30 core-seconds of CPU work
            rate, dt = cpu_port_rate(frames_cpu, thr)
            cpu = {"value": rate, "unit": UNIT, "cores": thr, "kind": "port",
                   "sample": f"{frames_cpu} frames x {PASSES} passes of the same workload, {dt:.1f} s wall",
                   **reference_numpy_rate()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": cfg2_config(F, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_source, "kernel": {1: "fused_unit_kernel<8>", 2: "fused_unit_kernel_v2<8>", 3: "fused_unit_kernel_v3<8>", 4: "fused_mma8_kernel"}[args.fused_impl], "bytes_per_px": BYTES_PER_PX,
                         "px_per_launch": px_launch, "launch_ms": launch_s * 1e3, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "note": "peak is the driver's b.copy_(a) probe (1 read : 1 write); this write-dominated "
                                 "stream can run marginally faster than that probe, so frac may exceed 1"},
            "cpu_baseline": cpu, "e2e": e2e, "e2e_noise": e2e_noise, "gpu_launches": launches, "clocks": clocks,
            "secondary": secondary,
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
    ap.add_argument("--e2e-frames", type=int, default=128, help="frames per pass of the e2e leg (default: the full config)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg3 / cfg5 / cfg4 measurements")
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
