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


def frame_stats(src: torch.Tensor, result) -> torch.Tensor:
    """(sse, n_samples, sum_sad_cost, nonzero_levels) of one coded frame as an int64 device tensor.
    sse / n_samples follow metrics.py:7-21 over the WHOLE plane (uncovered rows count, as in
    __main__.py:135-137)."""
    from . import batched
    ss = batched.sse_sad(src, result.recon_plane)
    nnz = batched.count_nonzero_batched(result.levels) if result.levels is not None else ss.new_zeros(())
    cost = result.costs.sum(dtype=torch.int64) if result.costs is not None else ss.new_zeros(())
    n = torch.full((), src.numel(), dtype=torch.int64, device=ss.device)
    return torch.stack([ss[0], n, cost.to(torch.int64), nnz.to(torch.int64)])


def _stack_local_frames(frames, lo, hi, device):
    """The local frames [lo, hi) as one (F, H, W) int16 tensor on ``device``.  Host frames (numpy arrays or
    CPU tensors) are uploaded on the CURRENT stream into memory this function owns until the coder that
    reads it -- enqueued on the same stream -- has been ordered behind the copies: no side streams, no
    buffer that can be recycled while a kernel still reads it."""
    local = []
    for i in range(lo, hi):
        f = frames[i]
        if not isinstance(f, torch.Tensor):
            f = torch.from_numpy(np.ascontiguousarray(f, dtype=np.int16))
        local.append(f)
    if not local:
        return None
    shapes = {tuple(f.shape) for f in local}
    if len(shapes) != 1 or len(next(iter(shapes))) != 2:
        raise ValueError(f"frames of one call must share one (H, W) shape, got {sorted(shapes)}")
    H, W = local[0].shape
    out = torch.empty((len(local), H, W), dtype=torch.int16, device=device)
    for k, f in enumerate(local):
        out[k].copy_(f.to(torch.int16) if f.dtype != torch.int16 else f, non_blocking=True)
    return out


def encode_frames_sharded(frames: Sequence, size: int, cost: str = "sad", qp: int = 27,
                          recon_neighbours: bool = True, bit_depth: int = 8, group=None,
                          device: torch.device | None = None,
                          encode_fn: Callable | None = None, stats_fn: Callable | None = None):
    """Code ``frames`` (a sequence of (H, W) int16 arrays / tensors, identical on every rank) with the
    contiguous shard ``shard_range(len(frames), rank, world)`` on this rank, then all-gather the per-frame
    statistics.

    The local frames go through ONE ``nh_encode_frames`` call (batched.encode_frames): their block rows
    share the wavefront scheduler, so several frames per GPU fill it; the per-frame statistics come from
    the same call.  Returns ``(local_results, stats, psnr)``: the FrameResult objects of this rank's frames
    (views into the batched outputs), an (n_frames, 4) int64 tensor [sse, n, sum_cost, nnz] identical on
    every rank, and the per-frame PSNR list (float64, finished on the host from the exact integer SSE).
    ``encode_fn`` / ``stats_fn`` (per-frame callables) exist so the host-side logic can be exercised on
    CPU (gloo)."""
    from . import batched
    rank, world = _world(group)
    n = len(frames)
    lo, hi = shard_range(n, rank, world)
    local, local_stats = [], []
    if encode_fn is None:
        dev = torch.device(device) if device is not None else None
        if dev is None:
            for i in range(lo, hi):
                if isinstance(frames[i], torch.Tensor) and frames[i].is_cuda:
                    dev = frames[i].device
                    break
        if hi > lo:
            if dev is None or dev.type != "cuda":
                raise RuntimeError("encode_frames_sharded needs a CUDA device (there is no CPU fallback)")
            with torch.cuda.device(dev):
                planes = _stack_local_frames(frames, lo, hi, dev)
                r = batched.encode_frames(planes, size, cost=cost, qp=qp, recon_neighbours=recon_neighbours,
                                          bit_depth=bit_depth)
            for k in range(hi - lo):
                local.append(batched.FrameResult(r.modes[k], r.costs[k], r.pred[k], r.coeff[k], r.levels[k],
                                                 r.recon_planes[k]))
                local_stats.append(r.stats[k])
    else:
        for i in range(lo, hi):
            f = frames[i]
            if not isinstance(f, torch.Tensor):
                f = torch.from_numpy(np.ascontiguousarray(f, dtype=np.int16))
            if device is not None:
                f = f.to(device)
            r = encode_fn(f)
            local.append(r)
            local_stats.append((stats_fn or frame_stats)(f, r).to(torch.int64))
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
