"""Multi-GPU driver: frames (or block ranges) are partitioned across ranks with NO collective on
the hot path; ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in CPU tests) is used
only for the final gather of per-frame statistics and, optionally, of the levels.

One process per GPU (``torchrun``); rank r owns the contiguous slice ``shard_range(n, r, world)``.
Intra-frame splitting is not done for the wavefront coder (config 5): its dependencies would need
a per-wave halo exchange, so a frame is the smallest unit ("replicas only" inside a frame).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first ``n % world`` ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n_items), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


_STREAMS: dict = {}


def _frame_streams(device, n):
    """Per-device side streams, created once: the caching allocator keeps one pool per stream, so
    fresh streams on every call would turn every output allocation into a cudaMalloc."""
    key = str(torch.device(device))
    have = _STREAMS.setdefault(key, [])
    while len(have) < n:
        have.append(torch.cuda.Stream(device=device))
    return have[:n]


def frame_stats(src: torch.Tensor, result) -> torch.Tensor:
    """(sse, n_samples, sum_sad_cost, nonzero_levels) of one coded frame as an int64 device tensor.
    sse / n_samples follow metrics.py:7-21 over the WHOLE plane (uncovered rows count, as in
    __main__.py:135-137)."""
    from . import batched
    ss = batched.sse_sad(src, result.recon_plane)
    nnz = batched.count_nonzero_batched(result.levels) if result.levels is not None else ss.new_zeros(())
    cost = result.costs.sum(dtype=torch.int64) if result.costs is not None else ss.new_zeros(())
    # torch.full, not new_tensor: a host scalar would be copied with a stream synchronisation, which
    # serialises frames that are meant to overlap on separate streams
    n = torch.full((), src.numel(), dtype=torch.int64, device=ss.device)
    return torch.stack([ss[0], n, cost.to(torch.int64), nnz.to(torch.int64)])


def encode_frames_sharded(frames: Sequence, size: int, cost: str = "sad", qp: int = 27,
                          recon_neighbours: bool = True, bit_depth: int = 8, group=None,
                          device: torch.device | None = None,
                          encode_fn: Callable | None = None, stats_fn: Callable | None = None,
                          max_concurrent_frames: int = 8):
    """Code ``frames`` (a sequence of (H, W) int16 arrays / tensors, identical on every rank) with
    frame i on rank ``i mod``-contiguous shard, then all-gather the per-frame statistics.

    Returns ``(local_results, stats, psnr)``: the FrameResult objects of this rank's frames, an
    (n_frames, 4) int64 tensor [sse, n, sum_cost, nnz] identical on every rank, and the per-frame
    PSNR list (float64, finished on the host from the exact integer SSE).
    ``encode_fn`` / ``stats_fn`` exist so the host-side logic can be exercised on CPU (gloo)."""
    from . import batched
    rank, world = _world(group)
    n = len(frames)
    lo, hi = shard_range(n, rank, world)
    if encode_fn is None:
        encode_fn = lambda f: batched.encode_frame(f, size, cost=cost, qp=qp, recon_neighbours=recon_neighbours,
                                                   bit_depth=bit_depth)
        stats_fn = frame_stats
    local, local_stats = [], []
    # A wavefront frame occupies only a few hundred warps (one per block row), so the frames of this
    # rank run concurrently on separate CUDA streams; the coders are independent (no shared state).
    use_streams = (device is not None and torch.device(device).type == "cuda" and hi - lo > 1
                   and max_concurrent_frames > 1)
    streams = _frame_streams(device, min(hi - lo, max_concurrent_frames)) if use_streams else []
    main = torch.cuda.current_stream(device) if use_streams else None
    for n_done, i in enumerate(range(lo, hi)):
        f = frames[i]
        if not isinstance(f, torch.Tensor):
            f = torch.from_numpy(np.ascontiguousarray(f, dtype=np.int16))
        if device is not None:
            f = f.to(device, non_blocking=True)
        if use_streams:
            st = streams[n_done % len(streams)]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                r = encode_fn(f)
                local_stats.append(stats_fn(f, r).to(torch.int64))
            local.append(r)
        else:
            r = encode_fn(f)
            local.append(r)
            local_stats.append(stats_fn(f, r).to(torch.int64))
    for st in streams:
        main.wait_stream(st)
    stat_dev = local_stats[0].device if local_stats else (device or torch.device("cpu"))
    # pad every rank's block to the largest shard so a single all_gather suffices
    per = -(-n // world) if world else n
    buf = torch.full((per, 5), -1, dtype=torch.int64, device=stat_dev)
    for k, s in enumerate(local_stats):
        buf[k, 0] = lo + k
        buf[k, 1:] = s
    if world > 1:
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf, group=group)
        allbuf = torch.cat(gathered).cpu()
    else:
        allbuf = buf.cpu()
    stats = torch.zeros((n, 4), dtype=torch.int64)
    for row in allbuf.tolist():
        if row[0] >= 0:
            stats[row[0]] = torch.tensor(row[1:], dtype=torch.int64)
    psnr = [batched.psnr_from_sse(int(s[0]), int(s[1])) for s in stats]
    return local, stats, psnr


def gather_levels(levels: torch.Tensor, dst: int = 0, group=None):
    """Final gather of per-GPU level tensors (equal shapes) onto ``dst``; returns the list there,
    None elsewhere."""
    rank, world = _world(group)
    if world == 1:
        return [levels]
    out = [torch.empty_like(levels) for _ in range(world)] if rank == dst else None
    dist.gather(levels, out, dst=dst, group=group)
    return out
