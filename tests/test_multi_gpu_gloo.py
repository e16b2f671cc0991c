"""Host-side logic of the multi-GPU path (frame sharding + final gather of statistics / levels)
exercised on CPU: two processes, gloo backend, the frame coder replaced by the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nano_hevc_b200.multi_gpu import shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 100):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from collections import namedtuple
    from nano_hevc_b200 import multi_gpu

    R = namedtuple("R", "levels recon_plane costs")
    rng = np.random.default_rng(5)
    frames = [rng.integers(0, 256, (24, 40)).astype(np.int16) for _ in range(5)]

    def encode(f):
        o = O.encode_frame(f.numpy(), 8, cost="sad", qp=27, recon_neighbours=True)
        return R(torch.from_numpy(o["levels"]), torch.from_numpy(o["recon_plane"]), torch.from_numpy(o["costs"]))

    def stats(f, r):
        sse = O.sse(f.numpy(), r.recon_plane.numpy())
        return torch.tensor([sse, f.numel(), int(r.costs.sum()), int(np.count_nonzero(r.levels.numpy()))])

    local, st, psnr = multi_gpu.encode_frames_sharded(frames, 8, encode_fn=encode, stats_fn=stats)
    lo, hi = multi_gpu.shard_range(len(frames), rank, world)
    assert len(local) == hi - lo
    lv = multi_gpu.gather_levels(torch.full((3,), rank, dtype=torch.int32), dst=0)
    q.put((rank, st.tolist(), psnr, None if lv is None else [int(t[0]) for t in lv]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharded_encode_and_gather():
    import oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    # every rank sees the statistics of ALL frames, and they equal a single-process run
    rng = np.random.default_rng(5)
    frames = [rng.integers(0, 256, (24, 40)).astype(np.int16) for _ in range(5)]
    want = []
    for f in frames:
        o = O.encode_frame(f, 8, cost="sad", qp=27, recon_neighbours=True)
        want.append([O.sse(f, o["recon_plane"]), f.size, int(o["costs"].sum()), int(np.count_nonzero(o["levels"]))])
    for rank, st, psnr, lv in res:
        assert st == want
        for p, f, w in zip(psnr, frames, want):
            assert p == pytest.approx(10 * np.log10(255 ** 2 / (w[0] / w[1])), rel=1e-12)
    assert res[0][3] == [0, 1] and res[1][3] is None
