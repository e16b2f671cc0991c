"""Device-side frame containers (nano_hevc_b200/frames.py) against goldens
produced by the reference's PackedFrame / FrameBufferPool (tests/golden/make_golden.py::g_containers)."""
import numpy as np
import pytest
import torch

from conftest import golden

SHAPES = ((6, 10), (18, 34), (64, 96))


def _run_pool_script(pool):
    """The scripted sequence of g_containers(); returns (trace, kept, cleared, messages)."""
    trace, msgs = [], {}
    snap = lambda kind, v: trace.append([kind, v, pool.available_count, pool.in_use_count])
    a, _ = pool.acquire(); snap(0, a)
    b, fb = pool.acquire(); snap(0, b)
    fb.y[:] = 7
    pool.release(a); snap(1, a)
    c, _ = pool.acquire(); snap(0, c)
    d, _ = pool.acquire(); snap(0, d)
    with pytest.raises(RuntimeError) as ei:
        pool.acquire()
    msgs["exhausted"] = str(ei.value)
    with pytest.raises(ValueError) as ei:
        pool.release(a + 100)
    msgs["release"] = str(ei.value)
    pool.release(b); snap(1, b)
    e2, fe = pool.acquire(clear=False); snap(0, e2)
    kept = int(fe.y[0, 0])
    pool.release(e2); snap(1, e2)
    _, fz = pool.acquire(clear=True)
    return trace, kept, int(fz.y[0, 0]), msgs


def test_pool_bookkeeping_matches_reference_cpu_tensors():
    """acquire / release order, counters and error messages of frame.py:224-293 (container logic
    only: CPU tensors, no kernel involved)."""
    from nano_hevc_b200.frames import DeviceFramePool, DevicePackedFrame
    g = golden("containers.npz")
    pool = DeviceFramePool(8, 8, pool_size=3, device="cpu")
    trace, kept, cleared, msgs = _run_pool_script(pool)
    assert trace == g["pool_trace"].tolist()
    assert kept == int(g["pool_kept_value"]) and cleared == int(g["pool_cleared_value"])
    assert msgs["exhausted"] == str(g["pool_exhausted_msg"]) and msgs["release"] == str(g["pool_release_msg"])
    assert pool.pool_size == 3
    # one arena: frames are consecutive slices of a single allocation, planes are views of the frame
    f0, f1 = pool._pool[0], pool._pool[1]
    assert f1.buffer.data_ptr() - f0.buffer.data_ptr() == f0.buffer.numel() * 2
    assert f0.u.data_ptr() == f0.buffer.data_ptr() + 64 * 2 and f0.v.data_ptr() == f0.u.data_ptr() + 16 * 2
    fr = DevicePackedFrame(6, 10, device="cpu")
    assert fr.y.shape == (6, 10) and fr.u.shape == (3, 5) and fr.v.shape == (3, 5) and fr.dtype == torch.int16
    fr.y[2, 3] = 5
    assert int(fr.buffer[2 * 10 + 3]) == 5
    fr.clear()
    assert not fr.buffer.any()


def test_conversion_needs_the_gpu():
    from nano_hevc_b200.frames import DevicePackedFrame
    fr = DevicePackedFrame(6, 10, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fr.to_yuv420p()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fr.load_yuv420p(bytes(90))
    with pytest.raises(ValueError):
        DevicePackedFrame(6, 10, dtype=torch.uint8, device="cpu").load_yuv420p(bytes(10))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_packed_frame_yuv420p_golden(shape):
    from nano_hevc_b200.frames import DevicePackedFrame
    g = golden("containers.npz")
    H, W = shape
    tag = f"{H}x{W}"
    raw = g[f"raw_{tag}"].tobytes()
    for dt in (torch.int16, torch.uint8):
        fr = DevicePackedFrame.from_yuv420p(raw, H, W, dtype=dt)
        for name in "yuv":
            assert np.array_equal(getattr(fr, name).cpu().numpy().astype(np.int64), g[f"{name}_{tag}"].astype(np.int64))
        assert fr.to_yuv420p() == raw
    # int16 samples outside 0..255 keep their low 8 bits, like astype(np.uint8)
    fr = DevicePackedFrame(H, W)
    fr.buffer.copy_(torch.from_numpy(g[f"i16_{tag}"]))
    assert np.array_equal(np.frombuffer(fr.to_yuv420p(), np.uint8), g[f"i16_bytes_{tag}"])
    # from_planes copies into one allocation
    fr2 = DevicePackedFrame.from_planes(fr.y.clone(), fr.u.clone(), fr.v.clone())
    assert torch.equal(fr2.buffer, fr.buffer)


@pytest.mark.gpu
def test_pool_on_device_and_large_round_trip():
    from nano_hevc_b200.frames import DeviceFramePool
    g = golden("containers.npz")
    pool = DeviceFramePool(8, 8, pool_size=3)
    trace, kept, cleared, _ = _run_pool_script(pool)
    assert trace == g["pool_trace"].tolist() and kept == 7 and cleared == 0
    # a 4K frame through a pooled buffer: unaligned tail (2160*3840*1.5 is a multiple of 16, so use
    # an odd geometry as well)
    for (H, W) in ((2160, 3840), (1082, 1922)):
        rng = np.random.default_rng(H)
        raw = rng.integers(0, 256, H * W + 2 * (H // 2) * (W // 2), dtype=np.uint8)
        p = DeviceFramePool(H, W, pool_size=2)
        idx, fr = p.acquire()
        fr.load_yuv420p(raw.tobytes())
        assert np.array_equal(fr.y.cpu().numpy(), raw[:H * W].reshape(H, W).astype(np.int16))
        assert fr.to_yuv420p() == raw.tobytes()
        p.release(idx)
        idx2, fr2 = p.acquire(clear=True)
        assert idx2 == idx and not fr2.buffer.any()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,bs", [("64x64_8", 8), ("96x128_16", 16)])
def test_cli_encode_from_packed_frame(tag, bs):
    """The CLI-faithful coder fed from a device container: YUV420p bytes in, YUV420p bytes out."""
    from nano_hevc_b200 import frame_encode
    from nano_hevc_b200.frames import DevicePackedFrame
    g = golden("cli.npz")
    y = g[f"y_{tag}"]
    H, W = y.shape
    u = np.full((H // 2, W // 2), 128, np.uint8)
    raw = y.astype(np.uint8).tobytes() + u.tobytes() + u.tobytes()
    fr = DevicePackedFrame.from_yuv420p(raw, H, W)
    (ry, ru, rv), stats = frame_encode.encode_frame_intra(fr.y, fr.u, fr.v, bs)
    assert np.array_equal(ry.cpu().numpy(), g[f"recon_y_{tag}"])
    out = DevicePackedFrame.from_planes(ry, ru, rv).to_yuv420p()
    assert out[:H * W] == g[f"recon_y_{tag}"].astype(np.uint8).tobytes()
