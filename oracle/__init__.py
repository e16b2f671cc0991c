"""CPU oracle for nano-hevc's block-coding hot path (TEST INFRASTRUCTURE ONLY).

ctypes front-end of ``oracle/nh_oracle.c`` -- a plain-C restatement of the
reference's numpy functions (each C function cites the reference file:line it
follows).  Parity status: **pinned** against golden vectors generated from the
imported reference (``tests/golden/make_golden.py``) and the reference tests'
known-answer values; see ``tests/test_oracle_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  ``nano_hevc_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnh_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/nh_oracle.c -> oracle/_build/libnh_oracle.so (gcc, seconds)."""
    src = os.path.join(_HERE, "nh_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.nho_sad.restype = C.c_int64
        _lib.nho_satd_4x4.restype = C.c_int64
        _lib.nho_satd_block.restype = C.c_int64
        _lib.nho_residual_energy.restype = C.c_int64
        _lib.nho_sse.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i16(a):
    return np.ascontiguousarray(a, dtype=np.int16)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def n_host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def _split(n, parts):
    parts = max(1, min(parts, n))
    edges = np.linspace(0, n, parts + 1).astype(np.int64)
    return [(int(edges[i]), int(edges[i + 1])) for i in range(parts) if edges[i + 1] > edges[i]]


def _run_threads(fn, n, threads):
    chunks = _split(n, threads)
    if len(chunks) <= 1:
        for a, b in chunks:
            fn(a, b)
        return
    with ThreadPoolExecutor(len(chunks)) as ex:
        list(ex.map(lambda ab: fn(*ab), chunks))


# ----------------------------------------------------------------- tables
def get_matrix(size: int, use_dst: bool = False) -> np.ndarray:
    out = np.empty((size, size), np.int32)
    if lib().nho_get_matrix(size, int(use_dst), _p(out)):
        raise ValueError(f"Unsupported transform size: {size}")
    return out


def intra_pred_angle(mode: int) -> int:
    return lib().nho_intra_pred_angle(mode)


# ------------------------------------------------------- per-block functions
def forward_transform(residual, use_dst=False):
    r = _i32(residual)
    n = r.shape[0]
    out = np.empty((n, n), np.int32)
    if lib().nho_forward_transform(_p(r), _p(out), n, int(use_dst)):
        raise ValueError(f"Unsupported transform size: {n}")
    return out


def inverse_transform(coeff, use_dst=False):
    c = _i32(coeff)
    n = c.shape[0]
    out = np.empty((n, n), np.int32)
    if lib().nho_inverse_transform(_p(c), _p(out), n, int(use_dst)):
        raise ValueError(f"Unsupported transform size: {n}")
    return out


def get_qp_params(qp):
    per, rem = C.c_int(), C.c_int()
    lib().nho_get_qp_params(int(qp), C.byref(per), C.byref(rem))
    return per.value, rem.value


def quantize(coeff, qp, size, is_intra=True):
    c = _i32(coeff)
    out = np.empty_like(c)
    if lib().nho_quantize(_p(c), _p(out), C.c_int64(c.size), int(qp), int(size), int(is_intra)):
        raise ValueError(f"Unsupported block size: {size}")
    return out


def dequantize(level, qp, size=None):
    lv = _i32(level)
    out = np.empty_like(lv)
    lib().nho_dequantize(_p(lv), _p(out), C.c_int64(lv.size), int(qp))
    return out


def quantize_block(coeff, qp, is_intra=True):
    return quantize(coeff, qp, np.asarray(coeff).shape[-1], is_intra)


def dequantize_block(level, qp):
    return dequantize(level, qp)


def estimate_bits(level):
    """quant.py:153-168 (numpy restatement: float64 sum, truncated)."""
    a = np.abs(np.asarray(level))
    return int(np.sum(np.log2(a + 1) + (a > 0) * 2))


def count_nonzero(level):
    """quant.py:171-173."""
    return int(np.count_nonzero(level))


def intra_dc_predict(top, left, size):
    out = np.empty((size, size), np.int16)
    lib().nho_intra_dc_predict(_p(_i16(top)), _p(_i16(left)), size, _p(out))
    return out


def intra_planar_predict(top, left, top_right, bottom_left, size):
    out = np.empty((size, size), np.int16)
    lib().nho_intra_planar_predict(_p(_i16(top)), _p(_i16(left)), int(top_right), int(bottom_left),
                                   size, _p(out))
    return out


def intra_angular_predict(top, left, top_left, mode, size):
    t, l = _i16(top), _i16(left)
    out = np.empty((size, size), np.int16)
    if lib().nho_intra_angular_predict(_p(t), t.size, _p(l), l.size, int(top_left), int(mode), size,
                                       _p(out)):
        raise IndexError("mode out of range")
    return out


def residual_block(orig, pred):
    o, p = _i16(orig), _i16(pred)
    out = np.empty_like(o)
    lib().nho_residual_block(_p(o), _p(p), _p(out), C.c_int64(o.size))
    return out


def reconstruct_block(pred, residual):
    p, r = _i16(pred), _i32(residual)
    out = np.empty_like(p)
    lib().nho_reconstruct_block(_p(p), _p(r), _p(out), C.c_int64(p.size))
    return out


def clip_to_pixel_range(block, bit_depth=8):
    b = _i16(block)
    out = np.empty_like(b)
    lib().nho_clip_to_pixel_range(_p(b), _p(out), C.c_int64(b.size), int(bit_depth))
    return out


def sad(a, b):
    a, b = _i16(a), _i16(b)
    return int(lib().nho_sad(_p(a), _p(b), C.c_int64(a.size)))


def satd_4x4(a, b):
    a, b = _i16(a).reshape(4, 4), _i16(b).reshape(4, 4)
    return int(lib().nho_satd_4x4(_p(a), _p(b)))


def satd_block(a, b):
    a, b = _i16(a), _i16(b)
    return int(lib().nho_satd_block(_p(a), _p(b), a.shape[0]))


def residual_energy(res):
    r = _i16(res)
    return int(lib().nho_residual_energy(_p(r), C.c_int64(r.size)))


def sse(a, b):
    a, b = _i16(a), _i16(b)
    return int(lib().nho_sse(_p(a), _p(b), C.c_int64(a.size)))


def mse(a, b):
    a = np.asarray(a)
    return float(sse(a, b)) / float(a.size)


def psnr(a, b, peak=255):
    """metrics.py:13-21 with the integer-SSE formulation."""
    err = mse(a, b)
    if err == 0:
        return float("inf")
    return 10 * np.log10(peak ** 2 / err)


# ---------------------------------------------------------------- frame level
def gather_refs(plane, x, y, size, T, L):
    """block.py:38-55 via the K1 gather oracle.  Returns (top, left, corner), unpadded."""
    pl = _i16(plane)
    H, W = pl.shape
    top = np.empty(2 * size + 1, np.int16)
    left = np.empty(2 * size + 1, np.int16)
    tl, ll, c = C.c_int(), C.c_int(), C.c_int()
    lib().nho_gather_refs(_p(pl), H, W, W, x, y, T, L, _p(top), C.byref(tl), _p(left),
                          C.byref(ll), C.byref(c))
    return top[: tl.value].copy(), left[: ll.value].copy(), c.value


def gather_refs_frame(plane, size, T, L):
    """Padded (B, 2N+1) top/left and (B,) corner for every full block, raster order."""
    pl = _i16(plane)
    H, W = pl.shape
    bw, bh = W // size, H // size
    top = np.empty((bw * bh, 2 * size + 1), np.int16)
    left = np.empty((bw * bh, 2 * size + 1), np.int16)
    corner = np.empty(bw * bh, np.int16)
    f = lib().nho_gather_refs_padded
    for b in range(bw * bh):
        bx, by = b % bw, b // bw
        f(_p(pl), H, W, W, bx * size, by * size, size, T, L, _p(top[b]), _p(left[b]),
          C.c_void_p(corner.ctypes.data + 2 * b))
    return top, left, corner


def blocks_from_plane(plane, size):
    """(H, W) -> (B, N, N) block-major, iterate_blocks order (block.py:68-74)."""
    pl = np.asarray(plane)
    H, W = pl.shape
    bh, bw = H // size, W // size
    return np.ascontiguousarray(
        pl[: bh * size, : bw * size].reshape(bh, size, bw, size).transpose(0, 2, 1, 3)
    ).reshape(bh * bw, size, size)


def predict_mode(top, left, corner, mode, size):
    out = np.empty((size, size), np.int16)
    lib().nho_predict_mode(_p(_i16(top)), _p(_i16(left)), int(corner), int(mode), size, _p(out))
    return out


def block_pipeline(orig, pred, qp, is_intra=True, use_dst=False, bit_depth=8):
    o, p = _i16(orig), _i16(pred)
    n = o.shape[0]
    coeff = np.empty((n, n), np.int32)
    levels = np.empty((n, n), np.int32)
    recon = np.empty((n, n), np.int16)
    lib().nho_block_pipeline(_p(o), _p(p), n, int(qp), int(is_intra), int(use_dst), bit_depth,
                             _p(coeff), _p(levels), _p(recon))
    return coeff, levels, recon


def pipeline_dcplanar_batch(orig, top, left, top_right, bottom_left, mode, qp, is_intra=True,
                            use_dst=False, bit_depth=8, threads=1):
    """Config-2 pipeline on (B,N,N) blocks.  mode: int (0 planar / 1 DC) or (B,) uint8 array."""
    o = _i16(orig)
    B, N, _ = o.shape
    t, l, tr, bl = _i16(top), _i16(left), _i16(top_right), _i16(bottom_left)
    marr = None if np.isscalar(mode) else np.ascontiguousarray(mode, np.uint8)
    m = int(mode) if marr is None else 0
    pred = np.empty((B, N, N), np.int16)
    coeff = np.empty((B, N, N), np.int32)
    levels = np.empty((B, N, N), np.int32)
    recon = np.empty((B, N, N), np.int16)
    f = lib().nho_pipeline_dcplanar_batch

    def run(a, b):
        f(_p(o[a:b]), _p(t[a:b]), _p(l[a:b]), _p(tr[a:b]), _p(bl[a:b]),
          _p(marr[a:b]) if marr is not None else None, m, C.c_int64(b - a), N, int(qp),
          int(is_intra), int(use_dst), bit_depth, _p(pred[a:b]), _p(coeff[a:b]), _p(levels[a:b]),
          _p(recon[a:b]))

    _run_threads(run, B, threads)
    return pred, coeff, levels, recon


def pipeline_modes_batch(orig, top, left, corner, mode, qp, is_intra=True, use_dst=False,
                         bit_depth=8, threads=1):
    """Pipeline with any of the 35 modes from padded (B, 2N+1) refs."""
    o = _i16(orig)
    B, N, _ = o.shape
    t, l, c = _i16(top), _i16(left), _i16(corner)
    marr = None if np.isscalar(mode) else np.ascontiguousarray(mode, np.uint8)
    m = int(mode) if marr is None else 0
    pred = np.empty((B, N, N), np.int16)
    coeff = np.empty((B, N, N), np.int32)
    levels = np.empty((B, N, N), np.int32)
    recon = np.empty((B, N, N), np.int16)
    f = lib().nho_pipeline_modes_batch

    def run(a, b):
        f(_p(o[a:b]), _p(t[a:b]), _p(l[a:b]), _p(c[a:b]),
          _p(marr[a:b]) if marr is not None else None, m, C.c_int64(b - a), N, int(qp),
          int(is_intra), int(use_dst), bit_depth, _p(pred[a:b]), _p(coeff[a:b]), _p(levels[a:b]),
          _p(recon[a:b]))

    _run_threads(run, B, threads)
    return pred, coeff, levels, recon


def search_block(orig, top, left, corner, cost="sad"):
    o = _i16(orig)
    N = o.shape[0]
    cost_c = C.c_int32()
    pred = np.empty((N, N), np.int16)
    costs = np.zeros(35, np.int32)
    m = lib().nho_search_block(_p(o), _p(_i16(top)), _p(_i16(left)), int(corner), N,
                               int(cost == "satd"), C.byref(cost_c), _p(pred), _p(costs))
    return m, cost_c.value, pred, costs


def encode_frame(plane, size, cost="sad", qp=27, recon_neighbours=False, bit_depth=8, threads=1):
    """Frame coder oracle.  Returns a dict of block-major outputs."""
    pl = _i16(plane)
    H, W = pl.shape
    B = (H // size) * (W // size)
    n = size
    out = dict(
        modes=np.empty(B, np.uint8), costs=np.empty(B, np.int32),
        pred=np.empty((B, n, n), np.int16), coeff=np.empty((B, n, n), np.int32),
        levels=np.empty((B, n, n), np.int32), recon=np.empty((B, n, n), np.int16),
        recon_plane=np.zeros((H, W), np.int16),
    )
    f = lib().nho_encode_frame

    def run(a, b):
        f(_p(pl), H, W, size, int(cost == "satd"), int(qp), int(recon_neighbours), bit_depth,
          _p(out["modes"]), _p(out["costs"]), _p(out["pred"]), _p(out["coeff"]),
          _p(out["levels"]), _p(out["recon"]), _p(out["recon_plane"]), a, b)

    if recon_neighbours:
        run(0, B)
    else:
        _run_threads(run, B, threads)
    return out


def forward_transform_batch(res, use_dst=False, threads=1):
    r = _i32(res)
    B, N, _ = r.shape
    out = np.empty_like(r)
    f = lib().nho_forward_transform_batch
    _run_threads(lambda a, b: f(_p(r[a:b]), _p(out[a:b]), C.c_int64(b - a), N, int(use_dst)), B,
                 threads)
    return out


def inverse_transform_batch(coeff, use_dst=False, threads=1):
    c = _i32(coeff)
    B, N, _ = c.shape
    out = np.empty_like(c)
    f = lib().nho_inverse_transform_batch
    _run_threads(lambda a, b: f(_p(c[a:b]), _p(out[a:b]), C.c_int64(b - a), N, int(use_dst)), B,
                 threads)
    return out
