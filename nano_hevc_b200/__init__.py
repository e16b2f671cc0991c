"""nano_hevc_b200 -- B200-native (sm_100a) block-coding hot path of nano-hevc.

Drop-in for the hot-path names of ``nano_hevc/__init__.py:5-48`` (same positional / keyword
parameters, return dtypes and shapes), plus ``*_batched`` variants on PyTorch CUDA tensors
(``nano_hevc_b200.batched``) and frame-level entry points.  Every function runs on the GPU
through the C ABI of ``libnh_b200.so``; there is no CPU implementation in this package.

Per-block functions take and return numpy arrays exactly like the reference: each call is
the batched kernel with B = 1 plus the host<->device copies.
"""
from __future__ import annotations

import ctypes as _C

import numpy as np

from . import _lib

__version__ = "0.1.0"


# ------------------------------------------------------------------ constants
def _matrix(size, use_dst=False):
    out = np.empty((size, size), dtype=np.int32)
    _lib.check(_lib.lib().nh_get_transform_matrix(size, int(use_dst), out.ctypes.data_as(_C.c_void_p)))
    return out


def _angles():
    out, a = [], _C.c_int()
    for mode in range(2, 35):
        _lib.check(_lib.lib().nh_get_intra_pred_angle(mode, _C.byref(a)))
        out.append(a.value)
    return out


def _scales():
    q, d = [], []
    a, b = _C.c_int(), _C.c_int()
    for rem in range(6):
        _lib.check(_lib.lib().nh_get_quant_scales(rem, _C.byref(a), _C.byref(b)))
        q.append(a.value)
        d.append(b.value)
    return q, d


DST4 = _matrix(4, True)      # transform.py:20-25
DCT4 = _matrix(4)            # transform.py:28-33
DCT8 = _matrix(8)            # transform.py:35-44
DCT16 = _matrix(16)          # transform.py:46-63
DCT32 = _matrix(32)          # transform.py:65-135
INTRA_PRED_ANGLE = _angles()             # intra.py:24-29
QUANT_SCALE, DEQUANT_SCALE = _scales()   # quant.py:21-22


# -------------------------------------------------------------------- helpers
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("nano_hevc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _dev(a, dtype):
    """numpy array -> CUDA tensor of `dtype` (numpy does the reference's astype() wrap-around)."""
    torch = _torch()
    arr = np.ascontiguousarray(np.asarray(a).astype(dtype, copy=False))
    return torch.from_numpy(arr.copy() if not arr.flags.writeable else arr).cuda()


def _host(t):
    return t.cpu().numpy()


def _size_of(block, what):
    block = np.asarray(block)
    size = block.shape[0]
    if size not in (4, 8, 16, 32):
        raise ValueError(f"Unsupported transform size: {size}")  # transform.py:151
    if block.shape != (size, size):
        raise ValueError(f"{what} must be square, got {block.shape}")
    return block, size


# ----------------------------------------------------------------- transforms
def forward_transform(residual: np.ndarray, use_dst: bool = False) -> np.ndarray:
    """transform.py:154-196."""
    from . import batched
    residual, size = _size_of(residual, "residual")
    out = batched.forward_transform_batched(_dev(residual, np.int32).reshape(1, size, size), use_dst)
    return _host(out[0])


def inverse_transform(coeff: np.ndarray, use_dst: bool = False) -> np.ndarray:
    """transform.py:199-238."""
    from . import batched
    coeff, size = _size_of(coeff, "coeff")
    out = batched.inverse_transform_batched(_dev(coeff, np.int32).reshape(1, size, size), use_dst)
    return _host(out[0])


def forward_transform_4x4(residual, use_dst=False):  # transform.py:241-243
    return forward_transform(residual, use_dst)


def inverse_transform_4x4(coeff, use_dst=False):  # transform.py:246-248
    return inverse_transform(coeff, use_dst)


def forward_transform_8x8(residual):  # transform.py:251-253
    return forward_transform(residual, use_dst=False)


def inverse_transform_8x8(coeff):  # transform.py:256-258
    return inverse_transform(coeff, use_dst=False)


def forward_transform_16x16(residual):  # transform.py:261-263
    return forward_transform(residual, use_dst=False)


def inverse_transform_16x16(coeff):  # transform.py:266-268
    return inverse_transform(coeff, use_dst=False)


def forward_transform_32x32(residual):  # transform.py:271-273
    return forward_transform(residual, use_dst=False)


def inverse_transform_32x32(coeff):  # transform.py:276-278
    return inverse_transform(coeff, use_dst=False)


# ---------------------------------------------------------------------- quant
def get_qp_params(qp: int):
    """quant.py:25-38 (host arithmetic inside the library; no device needed)."""
    per, rem = _C.c_int(), _C.c_int()
    _lib.check(_lib.lib().nh_get_qp_params(int(qp), _C.byref(per), _C.byref(rem)))
    return per.value, rem.value


def quantize(coeff: np.ndarray, qp: int, size: int, is_intra: bool = True) -> np.ndarray:
    """quant.py:41-79."""
    from . import batched
    c = np.asarray(coeff)
    out = batched.quantize_batched(_dev(c, np.int32), qp, size=int(size), is_intra=is_intra)
    return _host(out).reshape(c.shape)


def dequantize(level: np.ndarray, qp: int, size: int) -> np.ndarray:
    """quant.py:82-123 (``size`` is ignored there too)."""
    from . import batched
    lv = np.asarray(level)
    out = batched.dequantize_batched(_dev(lv, np.int32), qp)
    return _host(out).reshape(lv.shape)


def quantize_block(coeff: np.ndarray, qp: int, is_intra: bool = True) -> np.ndarray:
    """quant.py:126-137."""
    return quantize(coeff, qp, np.asarray(coeff).shape[0], is_intra)


def dequantize_block(level: np.ndarray, qp: int) -> np.ndarray:
    """quant.py:140-150."""
    return dequantize(level, qp, np.asarray(level).shape[0])


def count_nonzero(levels: np.ndarray) -> int:
    """quant.py:171-173."""
    from . import batched
    return int(batched.count_nonzero_batched(_dev(levels, np.int32)).item())


def estimate_bits(level: np.ndarray) -> int:
    """quant.py:153-168."""
    from . import batched
    return batched.estimate_bits_batched(_dev(level, np.int32))


def is_all_zero(levels: np.ndarray) -> bool:
    """quant.py:176-178."""
    return count_nonzero(levels) == 0


# ---------------------------------------------------------------------- intra
def _check_pred_size(size):
    if size not in (4, 8, 16, 32):
        raise ValueError(f"Unsupported block size: {size}")


def intra_dc_predict(top: np.ndarray, left: np.ndarray, size: int) -> np.ndarray:
    """intra.py:46-62."""
    from . import batched
    _check_pred_size(size)
    t = _dev(np.asarray(top).reshape(-1), np.int16).reshape(1, -1)
    l = _dev(np.asarray(left).reshape(-1), np.int16).reshape(1, -1)
    if t.shape[1] == size and l.shape[1] == size:
        return _host(batched.intra_dc_predict_batched(t, l, size)[0])
    # intra.py:61 sums the whole arrays whatever their length
    torch = _torch()
    out = torch.empty((1, size, size), dtype=torch.int16, device=t.device)
    _lib.check(_lib.lib().nh_intra_dc_predict_ragged(t.data_ptr(), t.shape[1], l.data_ptr(), l.shape[1],
                                                     out.data_ptr(), 1, size,
                                                     torch.cuda.current_stream().cuda_stream))
    return _host(out[0])


def intra_dc_predict_4x4(top: np.ndarray, left: np.ndarray) -> np.ndarray:
    """intra.py:37-43 ((sum + 4) >> 3 == (sum + 4) // 8)."""
    return intra_dc_predict(top, left, 4)


def intra_planar_predict(top, left, top_right: int, bottom_left: int, size: int) -> np.ndarray:
    """intra.py:81-113."""
    from . import batched
    _check_pred_size(size)
    t = _dev(np.asarray(top).reshape(-1)[:size], np.int16).reshape(1, -1)
    l = _dev(np.asarray(left).reshape(-1)[:size], np.int16).reshape(1, -1)
    tr = _dev(np.array([int(top_right)]), np.int16)
    bl = _dev(np.array([int(bottom_left)]), np.int16)
    return _host(batched.intra_planar_predict_batched(t, l, tr, bl, size)[0])


def _equivalent_angular_mode(mode: int) -> int:
    """intra.py:142-143 index INTRA_PRED_ANGLE[mode - 2] with Python list semantics: modes
    below 2 wrap around (negative index) and are treated as horizontal (mode < 18)."""
    mode = int(mode)
    if mode > 34 or mode < -31:
        raise IndexError("list index out of range")
    if mode >= 2:
        return mode
    wrapped = mode + 33  # same angle as this mode ...
    if wrapped == 18:
        # angle -32 applied horizontally exists in no HEVC mode (documented deviation)
        raise ValueError("mode -15 (angle -32 along the left references) is not supported")
    return 36 - wrapped if wrapped > 18 else wrapped  # ... applied along the left references


def _pad_ref(a, size, replicate):
    """Bring a reference array to 2*size+1 entries: the primary side is padded by repeating its
    last element (intra.py:174-178); entries of the secondary side beyond its length are never
    copied into the projection (intra.py:184-186), i.e. they behave as zeros."""
    a = np.asarray(a).reshape(-1).astype(np.int16, copy=False)
    w = 2 * size + 1
    if a.size >= w:
        return a[:w]
    fill = a[-1] if replicate else 0
    return np.concatenate([a, np.full(w - a.size, fill, dtype=np.int16)])


def intra_angular_predict(top, left, top_left: int, mode: int, size: int) -> np.ndarray:
    """intra.py:116-156."""
    from . import batched
    _check_pred_size(size)
    m = _equivalent_angular_mode(mode)
    vertical = m >= 18
    t = _dev(_pad_ref(top, size, replicate=vertical), np.int16).reshape(1, -1)
    l = _dev(_pad_ref(left, size, replicate=not vertical), np.int16).reshape(1, -1)
    c = _dev(np.array([int(top_left)]).astype(np.int16), np.int16)
    return _host(batched.intra_angular_predict_batched(t, l, c, m, size)[0])


def residual_block(orig: np.ndarray, pred: np.ndarray) -> np.ndarray:
    """intra.py:65-67."""
    from . import batched
    o = np.asarray(orig)
    return _host(batched.residual_block_batched(_dev(o, np.int16), _dev(pred, np.int16))).reshape(o.shape)


def reconstruct_block(pred: np.ndarray, residual: np.ndarray) -> np.ndarray:
    """intra.py:70-72."""
    from . import batched
    p = np.asarray(pred)
    # astype(int16) of the residual first (intra.py:72), then the wrap-around add on the device
    r = np.asarray(residual).astype(np.int16).astype(np.int32)
    return _host(batched.reconstruct_block_batched(_dev(p, np.int16), _dev(r, np.int32))).reshape(p.shape)


def clip_to_pixel_range(block: np.ndarray, bit_depth: int = 8) -> np.ndarray:
    """intra.py:75-78."""
    from . import batched
    b = np.asarray(block)
    clipped = np.clip(b, -32768, 32767) if b.dtype.itemsize > 2 else b  # keep wide inputs in range
    return _host(batched.clip_to_pixel_range_batched(_dev(clipped, np.int16), bit_depth)).reshape(b.shape)


# -------------------------------------------------------------------- metrics
# The reference widens before it reduces (metrics.py:9 float64, :26 / :33 int32, :48 int64).  Inputs whose
# dtype fits int16 take the int16 kernels; anything wider (uint16 samples, the int32 output of
# inverse_transform, int64) goes through the int32 kernels after the same astype(int32) wrap-around as
# metrics.py:26, or -- for mse on values an int32 cannot hold -- through the float64 kernel.
_NARROW = (np.dtype(np.int8), np.dtype(np.uint8), np.dtype(np.int16), np.dtype(np.bool_))


def _is_narrow(*arrs):
    return all(a.dtype in _NARROW for a in arrs)


def _wide_reduce(a32, b32):
    """(sum d^2 mod 2^64, sum |d| int32-wrap, float64 sum d^2) of int32 arrays on the device."""
    torch = _torch()
    x = _dev(a32, np.int32).reshape(-1)
    y = _dev(b32, np.int32).reshape(-1) if b32 is not None else None
    out = torch.empty((2,), dtype=torch.int64, device=x.device)
    fs = torch.empty((1,), dtype=torch.float64, device=x.device)
    _lib.check(_lib.lib().nh_reduce_metrics_i32(x.data_ptr(), None if y is None else y.data_ptr(), x.numel(),
                                                out.data_ptr(), fs.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    o = out.cpu().numpy()
    return int(o[0]), int(o[1]), float(fs.item())


def _pair16(a, b):
    return _dev(a, np.int16).reshape(-1), _dev(b, np.int16).reshape(-1)


def sad(a: np.ndarray, b: np.ndarray) -> int:
    """metrics.py:24-26."""
    from . import batched
    a, b = np.asarray(a), np.asarray(b)
    if _is_narrow(a, b):
        x, y = _pair16(a, b)
        return int(batched.sse_sad(x, y)[1].item())
    return _wide_reduce(a.astype(np.int32), b.astype(np.int32))[1]


def _sum_sq(a, b):
    """Sum of squared differences as the float64 value np.sum(diff ** 2) has (metrics.py:9-10)."""
    from . import batched
    torch = _torch()
    if _is_narrow(a, b):
        x, y = _pair16(a, b)
        return np.float64(int(batched.sse_sad(x, y)[0].item()))
    fits32 = all(np.issubdtype(v.dtype, np.integer) and
                 (v.dtype.itemsize < 4 or v.dtype == np.int32 or
                  (v.size == 0 or (int(v.min()) >= -2**31 and int(v.max()) < 2**31))) for v in (a, b))
    if fits32:
        sse, _, fsum = _wide_reduce(a.astype(np.int32), b.astype(np.int32))
        # below 2^53 every square and every partial sum is an exactly representable integer, so the
        # float64 sum of the reference equals the integer sum whatever its order
        return np.float64(sse) if fsum < 2.0 ** 53 else np.float64(fsum)
    x = _dev(a, np.float64).reshape(-1)
    y = _dev(b, np.float64).reshape(-1)
    fs = torch.empty((1,), dtype=torch.float64, device=x.device)
    _lib.check(_lib.lib().nh_reduce_sse_f64(x.data_ptr(), y.data_ptr(), x.numel(), fs.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
    return np.float64(fs.item())


def mse(original: np.ndarray, reconstructed: np.ndarray) -> float:
    """metrics.py:7-10 (exact integer SSE on the device, float64 mean on the host)."""
    a, b = np.asarray(original), np.asarray(reconstructed)
    return float(_sum_sq(a, b) / np.float64(a.size))


def psnr(original: np.ndarray, reconstructed: np.ndarray, peak: int = 255) -> float:
    """metrics.py:13-21."""
    err = mse(original, reconstructed)
    if err == 0:
        return float("inf")
    return 10 * np.log10(peak ** 2 / err)


def satd_4x4(a: np.ndarray, b: np.ndarray) -> int:
    """metrics.py:29-43."""
    from . import batched
    a, b = np.asarray(a).reshape(4, 4), np.asarray(b).reshape(4, 4)
    if _is_narrow(a, b):
        x, y = _dev(a, np.int16).reshape(1, 4, 4), _dev(b, np.int16).reshape(1, 4, 4)
        return int(batched.block_costs(x, y, outputs=("satd",))[1][0].item())
    torch = _torch()
    x, y = _dev(a.astype(np.int32), np.int32), _dev(b.astype(np.int32), np.int32)
    out = torch.empty((1,), dtype=torch.int64, device=x.device)
    _lib.check(_lib.lib().nh_satd_4x4_i32(x.data_ptr(), y.data_ptr(), 1, out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    return int(out.item())


def residual_energy(residual: np.ndarray) -> int:
    """metrics.py:46-48."""
    from . import batched
    r = np.asarray(residual)
    if _is_narrow(r):
        x = _dev(r, np.int16).reshape(-1)
        return int(batched.sse_sad(x, _torch().zeros_like(x))[0].item())
    if r.dtype == np.int32 or r.dtype == np.uint16:
        return _wide_reduce(r.astype(np.int32), None)[0]
    raise ValueError(f"residual_energy: dtype {r.dtype} is outside the supported domain (int8 .. int32)")


__all__ = [
    "INTRA_PRED_ANGLE", "intra_dc_predict_4x4", "intra_dc_predict", "intra_planar_predict",
    "intra_angular_predict", "residual_block", "reconstruct_block", "clip_to_pixel_range",
    "forward_transform", "inverse_transform", "forward_transform_4x4", "inverse_transform_4x4",
    "forward_transform_8x8", "inverse_transform_8x8", "forward_transform_16x16",
    "inverse_transform_16x16", "forward_transform_32x32", "inverse_transform_32x32",
    "DCT4", "DCT8", "DCT16", "DCT32", "DST4", "quantize", "dequantize", "quantize_block",
    "dequantize_block", "get_qp_params", "count_nonzero", "is_all_zero", "estimate_bits", "QUANT_SCALE",
    "DEQUANT_SCALE", "psnr", "mse", "sad", "satd_4x4", "residual_energy",
]
