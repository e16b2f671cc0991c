"""Batched ``(B, N, N)`` variants of the reference's per-block functions on PyTorch CUDA tensors.

PyTorch only owns the device buffers and the stream; every function here is a thin
call into the C ABI of ``libnh_b200.so`` (``include/nh_b200.h``).  Dtype contract as in the
reference: pixels / predictions / residuals / reconstructions are int16,
coefficients / levels are int32.  Inputs are never modified; outputs are fresh tensors.
"""
from __future__ import annotations

from collections import namedtuple

import torch

from . import _lib

SIZES = (4, 8, 16, 32)

PipelineResult = namedtuple("PipelineResult", "pred coeff levels recon")
FrameResult = namedtuple("FrameResult", "modes costs pred coeff levels recon_plane")
FramesResult = namedtuple("FramesResult", "modes costs pred coeff levels recon_planes stats")


def _require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError(
                "nano_hevc_b200 batched ops need CUDA tensors (there is no CPU fallback); got "
                + (str(t.device) if isinstance(t, torch.Tensor) else type(t).__name__))
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors live on different devices: {dev} vs {t.device}")
    return dev


def _c(t, dtype):
    """Contiguous tensor of the given dtype (a copy only if needed)."""
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_size(size):
    if size not in SIZES:
        raise ValueError(f"Unsupported transform size: {size}")


def _blocks(t, name="blocks"):
    if t.dim() != 3 or t.shape[1] != t.shape[2]:
        raise ValueError(f"{name} must have shape (B, N, N), got {tuple(t.shape)}")
    _check_size(t.shape[1])
    return t.shape[0], t.shape[1]


# ----------------------------------------------------------------- transforms
def forward_transform_batched(residual: torch.Tensor, use_dst: bool = False) -> torch.Tensor:
    """transform.py:154-196 over (B,N,N); int16 or int32 residuals -> int32 coefficients."""
    dev = _require_cuda(residual)
    B, N = _blocks(residual, "residual")
    is32 = residual.dtype != torch.int16
    r = _c(residual, torch.int32 if is32 else torch.int16)
    out = torch.empty((B, N, N), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_forward_transform(_ptr(r), int(is32), _ptr(out), B, N, int(bool(use_dst)),
                                                   _stream()))
    return out


def inverse_transform_batched(coeff: torch.Tensor, use_dst: bool = False) -> torch.Tensor:
    """transform.py:199-238 over (B,N,N) int32."""
    dev = _require_cuda(coeff)
    B, N = _blocks(coeff, "coeff")
    c = _c(coeff, torch.int32)
    out = torch.empty((B, N, N), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_inverse_transform(_ptr(c), _ptr(out), B, N, int(bool(use_dst)), _stream()))
    return out


# ---------------------------------------------------------------------- quant
def quantize_batched(coeff: torch.Tensor, qp: int, size: int | None = None, is_intra: bool = True):
    """quant.py:41-79 element-wise; ``size`` defaults to the last dimension (quantize_block)."""
    dev = _require_cuda(coeff)
    size = int(coeff.shape[-1]) if size is None else int(size)
    if not 1 <= size <= 63:   # quant.py:72 takes int(log2(size)) of anything; the kernel covers 1..63
        raise ValueError(f"Unsupported block size: {size}")
    c = _c(coeff, torch.int32)
    out = torch.empty_like(c)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_quantize(_ptr(c), _ptr(out), c.numel(), int(qp), size, int(bool(is_intra)),
                                          _stream()))
    return out


def dequantize_batched(level: torch.Tensor, qp: int, size: int | None = None):
    """quant.py:82-123 element-wise (``size`` is ignored, as in the reference)."""
    dev = _require_cuda(level)
    lv = _c(level, torch.int32)
    out = torch.empty_like(lv)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_dequantize(_ptr(lv), _ptr(out), lv.numel(), int(qp), 4, _stream()))
    return out


# ----------------------------------------------------------------- predictors
def intra_dc_predict_batched(top: torch.Tensor, left: torch.Tensor, size: int) -> torch.Tensor:
    """intra.py:46-62; top, left (B, size) -> (B, size, size) int16."""
    dev = _require_cuda(top, left)
    _check_size(size)
    t, l = _c(top, torch.int16), _c(left, torch.int16)
    if t.shape != l.shape or t.dim() != 2 or t.shape[1] != size:
        raise ValueError(f"top/left must both be (B, {size}), got {tuple(t.shape)} / {tuple(l.shape)}")
    B = t.shape[0]
    out = torch.empty((B, size, size), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_intra_dc_predict(_ptr(t), _ptr(l), _ptr(out), B, size, _stream()))
    return out


def intra_planar_predict_batched(top, left, top_right, bottom_left, size: int) -> torch.Tensor:
    """intra.py:81-113; top, left (B, size); top_right, bottom_left (B,)."""
    dev = _require_cuda(top, left, top_right, bottom_left)
    _check_size(size)
    t, l = _c(top, torch.int16), _c(left, torch.int16)
    tr, bl = _c(top_right, torch.int16).reshape(-1), _c(bottom_left, torch.int16).reshape(-1)
    B = t.shape[0]
    if t.shape != (B, size) or l.shape != (B, size) or tr.numel() != B or bl.numel() != B:
        raise ValueError("planar refs must be top/left (B, size) and top_right/bottom_left (B,)")
    out = torch.empty((B, size, size), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_intra_planar_predict(_ptr(t), _ptr(l), _ptr(tr), _ptr(bl), _ptr(out), B,
                                                      size, _stream()))
    return out


def _mode_args(mode, B, dev, lo, hi=34):
    """(modes_tensor_or_None, scalar_mode) with the reference's range errors.  A per-block mode tensor is checked
    on the device (one min/max reduction and a host read) so that an out-of-range entry raises like the scalar
    path and the reference do, instead of being clamped by the kernel."""
    if isinstance(mode, torch.Tensor):
        if mode.numel() != B:
            raise ValueError(f"modes must have {B} entries, got {mode.numel()}")
        if B:
            mn, mx = (int(v) for v in torch.aminmax(mode.to(dev)))
            if mx > 34:
                raise IndexError("list index out of range")  # intra.py:142 indexes INTRA_PRED_ANGLE[mode - 2]
            if mn < lo or mx > hi:
                raise ValueError(f"modes tensor holds values outside {lo}..{hi} (min {mn}, max {mx})")
        return _c(mode.to(dev), torch.uint8).reshape(-1), 0
    mode = int(mode)
    if mode > 34:
        raise IndexError("list index out of range")  # intra.py:142 indexes INTRA_PRED_ANGLE[mode - 2]
    if mode < lo:
        raise ValueError(f"mode {mode} out of range {lo}..34")
    return None, mode


def _padded_refs(top, left, top_left, size):
    t, l, c = _c(top, torch.int16), _c(left, torch.int16), _c(top_left, torch.int16).reshape(-1)
    B = t.shape[0]
    w = 2 * size + 1
    if t.shape != (B, w) or l.shape != (B, w) or c.numel() != B:
        raise ValueError(f"angular refs must be top/left (B, {w}) and top_left (B,)")
    return t, l, c, B


def intra_angular_predict_batched(top, left, top_left, mode, size: int) -> torch.Tensor:
    """intra.py:116-207; top, left (B, 2*size+1) with index 0 = corner slot; top_left (B,);
    mode: int 2..34 or a (B,) uint8 tensor."""
    return intra_predict_modes_batched(top, left, top_left, mode, size, allow_dc_planar=False)


def intra_predict_modes_batched(top, left, top_left, mode, size: int, allow_dc_planar: bool = True):
    """Any of the 35 modes from padded (B, 2N+1) references (SURVEY.md 8a K1 convention:
    DC / planar use top[1:N+1], left[1:N+1], TR = top[N+1], BL = left[N+1])."""
    dev = _require_cuda(top, left, top_left)
    _check_size(size)
    t, l, c, B = _padded_refs(top, left, top_left, size)
    m, ms = _mode_args(mode, B, dev, 0 if allow_dc_planar else 2)
    out = torch.empty((B, size, size), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_intra_predict_modes(_ptr(t), _ptr(l), _ptr(c), _ptr(m), ms,
                                                     int(bool(allow_dc_planar)), _ptr(out), B, size,
                                                     _stream()))
    return out


# ------------------------------------------------- residual / recon / clip
def residual_block_batched(orig, pred):
    """intra.py:65-67 element-wise."""
    dev = _require_cuda(orig, pred)
    o, p = _c(orig, torch.int16), _c(pred, torch.int16)
    if o.shape != p.shape:
        raise ValueError("orig and pred must have the same shape")
    out = torch.empty_like(o)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_residual_block(_ptr(o), _ptr(p), _ptr(out), o.numel(), _stream()))
    return out


def reconstruct_block_batched(pred, residual):
    """intra.py:70-72 element-wise (int32 residual truncated to int16, wrap-around add)."""
    dev = _require_cuda(pred, residual)
    p, r = _c(pred, torch.int16), _c(residual, torch.int32)
    if p.shape != r.shape:
        raise ValueError("pred and residual must have the same shape")
    out = torch.empty_like(p)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_reconstruct_block(_ptr(p), _ptr(r), _ptr(out), p.numel(), _stream()))
    return out


def clip_to_pixel_range_batched(block, bit_depth: int = 8):
    """intra.py:75-78 element-wise."""
    dev = _require_cuda(block)
    b = _c(block, torch.int16)
    out = torch.empty_like(b)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_clip_to_pixel_range(_ptr(b), _ptr(out), b.numel(), int(bit_depth), _stream()))
    return out


# -------------------------------------------------------------- fused pipeline
def _outputs(want, B, N, dev):
    want = set(want)
    unknown = want - {"pred", "coeff", "levels", "recon"}
    if unknown:
        raise ValueError(f"unknown outputs {sorted(unknown)}")
    mk = lambda name, dt: torch.empty((B, N, N), dtype=dt, device=dev) if name in want else None
    return PipelineResult(mk("pred", torch.int16), mk("coeff", torch.int32), mk("levels", torch.int32),
                          mk("recon", torch.int16))


def fused_block_pipeline(orig, top, left, top_right, bottom_left, mode, qp: int, is_intra: bool = True,
                         use_dst: bool = False, bit_depth: int = 8,
                         outputs=("pred", "coeff", "levels", "recon"), out: PipelineResult | None = None):
    """K6: DC (mode 1) / planar (mode 0) predict -> residual -> forward -> quantize -> dequantize ->
    inverse -> reconstruct -> clip in one kernel (README.md:55-71 composition).

    orig (B,N,N) int16; top, left (B,N); top_right, bottom_left (B,); mode: int or (B,) uint8.
    ``out`` lets a caller reuse output tensors (benchmark loops)."""
    dev = _require_cuda(orig, top, left, top_right, bottom_left)
    B, N = _blocks(orig, "orig")
    o = _c(orig, torch.int16)
    t, l = _c(top, torch.int16), _c(left, torch.int16)
    tr, bl = _c(top_right, torch.int16).reshape(-1), _c(bottom_left, torch.int16).reshape(-1)
    if t.shape != (B, N) or l.shape != (B, N) or tr.numel() != B or bl.numel() != B:
        raise ValueError("refs must be top/left (B, N) and top_right/bottom_left (B,)")
    m, ms = _mode_args(mode, B, dev, 0, 1)
    if m is None and ms > 1:
        raise ValueError("fused_block_pipeline handles modes 0 (planar) and 1 (DC); use "
                         "fused_block_pipeline_modes for angular modes")
    res = out if out is not None else _outputs(outputs, B, N, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_fused_pipeline_dcplanar(
            _ptr(o), _ptr(t), _ptr(l), _ptr(tr), _ptr(bl), _ptr(m), ms, B, N, int(qp),
            int(bool(is_intra)), int(bool(use_dst)), int(bit_depth),
            _ptr(res.pred), _ptr(res.coeff), _ptr(res.levels), _ptr(res.recon), _stream()))
    return res


def fused_block_pipeline_modes(orig, top, left, top_left, mode, qp: int, is_intra: bool = True,
                               use_dst: bool = False, bit_depth: int = 8,
                               outputs=("pred", "coeff", "levels", "recon"),
                               out: PipelineResult | None = None):
    """K6 with any of the 35 modes from padded (B, 2N+1) references."""
    dev = _require_cuda(orig, top, left, top_left)
    B, N = _blocks(orig, "orig")
    o = _c(orig, torch.int16)
    t, l, c, B2 = _padded_refs(top, left, top_left, N)
    if B2 != B:
        raise ValueError("orig and refs disagree on the number of blocks")
    m, ms = _mode_args(mode, B, dev, 0)
    res = out if out is not None else _outputs(outputs, B, N, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_fused_pipeline_modes(
            _ptr(o), _ptr(t), _ptr(l), _ptr(c), _ptr(m), ms, B, N, int(qp), int(bool(is_intra)),
            int(bool(use_dst)), int(bit_depth),
            _ptr(res.pred), _ptr(res.coeff), _ptr(res.levels), _ptr(res.recon), _stream()))
    return res


def host_block_pipeline(orig, top, left, top_right, bottom_left, mode, qp: int, is_intra: bool = True,
                        use_dst: bool = False, bit_depth: int = 8,
                        outputs=("pred", "coeff", "levels", "recon"), chunk_blocks: int | None = None,
                        device: torch.device | None = None, scratch: torch.Tensor | None = None,
                        out: PipelineResult | None = None, int16_results: bool = False):
    """K6 on HOST buffers (numpy arrays or CPU tensors, pinned memory recommended) through
    ``nh_host_pipeline_dcplanar``: the library overlaps H2D, the kernel and D2H on internal streams
    and returns CPU tensors.  This is the call behind the ``e2e`` number of bench.py.
    ``int16_results``: coefficients and levels come back as int16 (``nh_host_pipeline_dcplanar_i16``:
    DMA straight into the result tensors, no host pass; raises when a block leaves the pixel domain)."""
    import numpy as np
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    as_cpu = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dt).contiguous()
    o = as_cpu(orig, torch.int16)
    B, N = _blocks(o, "orig")
    t, l = as_cpu(top, torch.int16), as_cpu(left, torch.int16)
    tr, bl = as_cpu(top_right, torch.int16).reshape(-1), as_cpu(bottom_left, torch.int16).reshape(-1)
    if t.shape != (B, N) or l.shape != (B, N) or tr.numel() != B or bl.numel() != B:
        raise ValueError("refs must be top/left (B, N) and top_right/bottom_left (B,)")
    for x in (o, t, l, tr, bl):
        if x.is_cuda:
            raise ValueError("host_block_pipeline takes host buffers; use fused_block_pipeline for CUDA tensors")
    if isinstance(mode, (torch.Tensor, np.ndarray)):
        m = as_cpu(mode, torch.uint8).reshape(-1)
        if m.numel() != B:
            raise ValueError(f"modes must have {B} entries")
        ms = 0
    else:
        m, ms = None, int(mode)
        if ms not in (0, 1):
            raise ValueError("mode must be 0 (planar) or 1 (DC)")
    if out is None:
        want = set(outputs)
        mk = lambda name, dt: torch.empty((B, N, N), dtype=dt).pin_memory() if name in want else None
        wide = torch.int16 if int16_results else torch.int32
        out = PipelineResult(mk("pred", torch.int16), mk("coeff", wide), mk("levels", wide), mk("recon", torch.int16))
    for res_t in (out.coeff, out.levels):
        if res_t is not None and res_t.dtype != (torch.int16 if int16_results else torch.int32):
            raise ValueError("coeff / levels of `out` must be int16 with int16_results, int32 otherwise")
    chunk = int(chunk_blocks) if chunk_blocks else max(1024, (1 << 23) // (N * N))
    L = _lib.lib()
    nbytes = int(L.nh_host_pipeline_scratch_bytes(N, chunk))
    if scratch is None or scratch.numel() < nbytes:
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        entry = L.nh_host_pipeline_dcplanar_i16 if int16_results else L.nh_host_pipeline_dcplanar
        _lib.check(entry(
            _ptr(o), _ptr(t), _ptr(l), _ptr(tr), _ptr(bl), _ptr(m), ms, B, N, int(qp), int(bool(is_intra)),
            int(bool(use_dst)), int(bit_depth), _ptr(out.pred), _ptr(out.coeff), _ptr(out.levels),
            _ptr(out.recon), _ptr(scratch), scratch.numel(), chunk))
    return out


# ------------------------------------------------------------------ frame level
def _plane(plane):
    if plane.dim() != 2:
        raise ValueError(f"plane must be (H, W), got {tuple(plane.shape)}")
    p = _c(plane, torch.int16)
    return p, p.shape[0], p.shape[1]


def gather_refs(plane, size: int, n_top: int | None = None, n_left: int | None = None):
    """K1: block.py:38-55 neighbours of every full block (iterate_blocks order) with the 128
    substitution at frame edges, padded to 2N+1 by replicate-last.  Returns (top, left, corner)."""
    dev = _require_cuda(plane)
    _check_size(size)
    p, H, W = _plane(plane)
    n_top = 2 * size if n_top is None else int(n_top)
    n_left = 2 * size if n_left is None else int(n_left)
    B = (H // size) * (W // size)
    top = torch.empty((B, 2 * size + 1), dtype=torch.int16, device=dev)
    left = torch.empty((B, 2 * size + 1), dtype=torch.int16, device=dev)
    corner = torch.empty((B,), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_gather_refs(_ptr(p), H, W, W, size, n_top, n_left, _ptr(top), _ptr(left),
                                             _ptr(corner), _stream()))
    return top, left, corner


def plane_to_blocks(plane, size: int):
    """(H, W) -> (B, N, N) in iterate_blocks order (block.py:68-74; partial blocks skipped)."""
    dev = _require_cuda(plane)
    _check_size(size)
    p, H, W = _plane(plane)
    B = (H // size) * (W // size)
    out = torch.empty((B, size, size), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_plane_to_blocks(_ptr(p), H, W, W, size, _ptr(out), _stream()))
    return out


def blocks_to_plane(blocks, height: int, width: int):
    """Inverse of plane_to_blocks; uncovered rows / columns are zero (frame.py:41-43)."""
    dev = _require_cuda(blocks)
    B, N = _blocks(blocks)
    b = _c(blocks, torch.int16)
    if B != (height // N) * (width // N):
        raise ValueError("block count does not match the plane size")
    out = torch.zeros((height, width), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_blocks_to_plane(_ptr(b), height, width, width, N, _ptr(out), _stream()))
    return out


def encode_frame(plane, size: int, cost: str = "sad", qp: int = 27, recon_neighbours: bool = False,
                 bit_depth: int = 8, outputs=("modes", "costs", "pred", "coeff", "levels")):
    """K7 / K8: exhaustive 35-mode search + winner pipeline over every full block of a plane.

    recon_neighbours=False: neighbours from the source plane, all blocks independent (config 3).
    recon_neighbours=True : neighbours from the reconstructed plane, anti-diagonal wavefront
    equivalent to the raster loop of block.py:68-74 (config 5).
    Returns FrameResult; recon_plane is always produced."""
    dev = _require_cuda(plane)
    _check_size(size)
    if cost not in ("sad", "satd"):
        raise ValueError("cost must be 'sad' or 'satd'")
    p, H, W = _plane(plane)
    B = (H // size) * (W // size)
    want = set(outputs)
    mk = lambda name, shape, dt: torch.empty(shape, dtype=dt, device=dev) if name in want else None
    res = FrameResult(mk("modes", (B,), torch.uint8), mk("costs", (B,), torch.int32),
                      mk("pred", (B, size, size), torch.int16), mk("coeff", (B, size, size), torch.int32),
                      mk("levels", (B, size, size), torch.int32),
                      torch.empty((H, W), dtype=torch.int16, device=dev))
    L = _lib.lib()
    nbytes = int(L.nh_encode_frame_scratch_bytes(H, W, size))
    scratch = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.nh_encode_frame(_ptr(p), H, W, W, size, int(cost == "satd"), int(qp),
                                     int(bool(recon_neighbours)), int(bit_depth), _ptr(res.modes),
                                     _ptr(res.costs), _ptr(res.pred), _ptr(res.coeff), _ptr(res.levels),
                                     _ptr(res.recon_plane), _ptr(scratch), scratch.numel(), _stream()))
    return res


def encode_frames(planes, size: int, cost: str = "sad", qp: int = 27, recon_neighbours: bool = False,
                  bit_depth: int = 8, outputs=("modes", "costs", "pred", "coeff", "levels"), stats: bool = True,
                  out: FramesResult | None = None, scratch: torch.Tensor | None = None):
    """K7 / K8 over a batch of frames in ONE call (``nh_encode_frames``): ``planes`` is an (F, H, W) int16
    tensor; frame f is coded exactly as ``encode_frame(planes[f], ...)`` would code it.  With
    ``recon_neighbours`` the block rows of all frames share the wavefront scheduler, so a batch fills the
    GPU (BASELINE config 5), without it every block of every frame is independent (config 3).

    Returns FramesResult with leading frame dimension: modes / costs (F, B), pred / coeff / levels
    (F, B, N, N), recon_planes (F, H, W) and, with ``stats``, an (F, 4) int64 device tensor
    [sse over the whole plane, H * W, sum of winner costs, non-zero levels] (metrics.py:7-21 numerators;
    ``psnr_from_sse`` finishes PSNR on the host).  ``out`` / ``scratch`` let a caller reuse buffers."""
    dev = _require_cuda(planes)
    _check_size(size)
    if cost not in ("sad", "satd"):
        raise ValueError("cost must be 'sad' or 'satd'")
    if planes.dim() != 3:
        raise ValueError(f"planes must be (F, H, W), got {tuple(planes.shape)}")
    p = _c(planes, torch.int16)
    F, H, W = p.shape
    B = (H // size) * (W // size)
    if out is None:
        want = set(outputs)
        if stats:
            want |= {"costs", "levels"}
        mk = lambda name, shape, dt: torch.empty(shape, dtype=dt, device=dev) if name in want else None
        out = FramesResult(mk("modes", (F, B), torch.uint8), mk("costs", (F, B), torch.int32),
                           mk("pred", (F, B, size, size), torch.int16), mk("coeff", (F, B, size, size), torch.int32),
                           mk("levels", (F, B, size, size), torch.int32),
                           torch.empty((F, H, W), dtype=torch.int16, device=dev),
                           torch.empty((F, 4), dtype=torch.int64, device=dev) if stats else None)
    L = _lib.lib()
    nbytes = int(L.nh_encode_frames_scratch_bytes(F, H, W, size)) if recon_neighbours else 16
    if scratch is None or scratch.numel() < nbytes:
        scratch = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.nh_encode_frames(_ptr(p), F, H * W, H, W, W, size, int(cost == "satd"), int(qp),
                                      int(bool(recon_neighbours)), int(bit_depth), _ptr(out.modes), _ptr(out.costs),
                                      _ptr(out.pred), _ptr(out.coeff), _ptr(out.levels), _ptr(out.recon_planes),
                                      _ptr(out.stats), _ptr(scratch), scratch.numel(), _stream()))
    return out


def host_encode_frames(planes, size: int, cost: str = "sad", qp: int = 27, recon_neighbours: bool = False,
                       bit_depth: int = 8, outputs=("modes", "costs", "pred", "coeff", "levels", "recon_planes"),
                       stats: bool = True, frames_per_chunk: int = 2, device: torch.device | None = None,
                       scratch: torch.Tensor | None = None, out: FramesResult | None = None):
    """``encode_frames`` on HOST buffers (numpy array or CPU tensor (F, H, W), pinned memory recommended)
    through ``nh_host_encode_frames``: chunks of ``frames_per_chunk`` frames rotate over three internal
    streams (upload, kernels, download overlap) and the results come back as CPU tensors in the
    reference's dtypes.  Only the ``outputs`` asked for cross PCIe.  This is the call behind the
    ``e2e`` numbers of BASELINE configs 3 and 5 in bench.py."""
    import numpy as np
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _check_size(size)
    if cost not in ("sad", "satd"):
        raise ValueError("cost must be 'sad' or 'satd'")
    p = planes if isinstance(planes, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(planes))
    if p.is_cuda:
        raise ValueError("host_encode_frames takes host buffers; use encode_frames for CUDA tensors")
    if p.dim() != 3:
        raise ValueError(f"planes must be (F, H, W), got {tuple(p.shape)}")
    p = p.to(torch.int16).contiguous()
    F, H, W = p.shape
    B = (H // size) * (W // size)
    if out is None:
        want = set(outputs)
        mk = lambda name, shape, dt: torch.empty(shape, dtype=dt).pin_memory() if name in want else None
        out = FramesResult(mk("modes", (F, B), torch.uint8), mk("costs", (F, B), torch.int32),
                           mk("pred", (F, B, size, size), torch.int16), mk("coeff", (F, B, size, size), torch.int32),
                           mk("levels", (F, B, size, size), torch.int32), mk("recon_planes", (F, H, W), torch.int16),
                           torch.empty((F, 4), dtype=torch.int64).pin_memory() if stats else None)
    L = _lib.lib()
    fc = max(1, min(int(frames_per_chunk), max(F, 1)))
    nbytes = int(L.nh_host_encode_frames_scratch_bytes(fc, H, W, size, int(bool(recon_neighbours))))
    if scratch is None or scratch.numel() < nbytes:
        scratch = torch.empty((max(nbytes, 256),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.nh_host_encode_frames(_ptr(p), F, H, W, size, int(cost == "satd"), int(qp),
                                           int(bool(recon_neighbours)), int(bit_depth), _ptr(out.modes), _ptr(out.costs),
                                           _ptr(out.pred), _ptr(out.coeff), _ptr(out.levels), _ptr(out.recon_planes),
                                           _ptr(out.stats), fc, _ptr(scratch), scratch.numel()))
    return out


# ------------------------------------------------------------------ reductions
def sse_sad(a, b):
    """Integer numerators of metrics.py mse / sad: returns a (2,) int64 device tensor [sse, sad]."""
    dev = _require_cuda(a, b)
    x, y = _c(a, torch.int16), _c(b, torch.int16)
    if x.shape != y.shape:
        raise ValueError("a and b must have the same shape")
    out = torch.empty((2,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_reduce_sse_sad(_ptr(x), _ptr(y), x.numel(), _ptr(out), _stream()))
    return out


def block_costs(a, b, outputs=("sad", "satd", "energy")):
    """Per-block SAD / SATD (sum of satd_4x4, metrics.py:29-43) / residual energy of (B,N,N) pairs."""
    dev = _require_cuda(a, b)
    B, N = _blocks(a)
    x, y = _c(a, torch.int16), _c(b, torch.int16)
    if x.shape != y.shape:
        raise ValueError("a and b must have the same shape")
    sad = torch.empty((B,), dtype=torch.int32, device=dev) if "sad" in outputs else None
    satd = torch.empty((B,), dtype=torch.int32, device=dev) if "satd" in outputs else None
    en = torch.empty((B,), dtype=torch.int64, device=dev) if "energy" in outputs else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_block_costs(_ptr(x), _ptr(y), B, N, _ptr(sad), _ptr(satd), _ptr(en), _stream()))
    return sad, satd, en


def count_nonzero_batched(levels):
    """quant.py:171-173 over a whole tensor: () int64 device tensor."""
    dev = _require_cuda(levels)
    lv = _c(levels, torch.int32)
    out = torch.empty((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_count_nonzero(_ptr(lv), lv.numel(), _ptr(out), _stream()))
    return out[0]


def estimate_bits_batched(levels) -> int:
    """quant.py:153-168 over a whole tensor: int(sum(log2(|l| + 1) + 2 * (l != 0)))."""
    dev = _require_cuda(levels)
    lv = _c(levels, torch.int32)
    nnz = torch.empty((1,), dtype=torch.int64, device=dev)
    s = torch.empty((1,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nh_level_stats(_ptr(lv), lv.numel(), _ptr(nnz), _ptr(s), _stream()))
    return int(float(s.item()) + 2.0 * int(nnz.item()))


def psnr_from_sse(sse: int, count: int, peak: int = 255) -> float:
    """metrics.py:13-21 finished on the host in float64 from the exact integer SSE."""
    import numpy as np
    if count == 0:
        return float("nan")
    err = float(np.float64(int(sse)) / np.float64(int(count)))
    if err == 0:
        return float("inf")
    return float(10 * np.log10(peak ** 2 / err))
