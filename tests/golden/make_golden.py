#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (Luodian/nano-hevc).

Run in the authoring container only (the reference is mounted read-only at
/root/reference and does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every array below is produced by calling the reference's own numpy functions
(nano_hevc.intra / transform / quant / metrics / block) on seeded inputs.  The
frame-level compositions (gather K1, search K7, raster recon-neighbour coder
K8) are literal Python loops over reference functions, exactly as SURVEY.md
section 8a defines them.  The committed .npz files are what tests/ compare the
C oracle and the CUDA path against.
"""
import os
import sys

import numpy as np

REF = os.environ.get("NANO_HEVC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import nano_hevc as R  # noqa: E402
from nano_hevc.quant import get_qp_params  # noqa: E402
from nano_hevc.block import BlockView, iterate_blocks  # noqa: E402
from nano_hevc.frame import Plane  # noqa: E402
from nano_hevc.__main__ import create_test_frame, encode_frame_intra  # noqa: E402
from nano_hevc import metrics as RM  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SIZES = (4, 8, 16, 32)


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print(name, {k: v.shape for k, v in arrs.items()})


# ------------------------------------------------------------------ tables
def g_tables():
    save("tables.npz", DCT4=R.DCT4, DCT8=R.DCT8, DCT16=R.DCT16, DCT32=R.DCT32, DST4=R.DST4,
         INTRA_PRED_ANGLE=np.array(R.INTRA_PRED_ANGLE, np.int32),
         QUANT_SCALE=np.array(R.QUANT_SCALE, np.int32),
         DEQUANT_SCALE=np.array(R.DEQUANT_SCALE, np.int32),
         qp_params=np.array([get_qp_params(q) for q in range(-3, 56)], np.int32))


# --------------------------------------------------------------- predictors
def g_predictors():
    rng = np.random.default_rng(2024)
    out = {}
    for N in SIZES:
        K = 6
        top = rng.integers(0, 256, (K, N)).astype(np.int16)
        left = rng.integers(0, 256, (K, N)).astype(np.int16)
        top[K - 1] = rng.integers(0, 1024, N)  # 10-bit case
        left[K - 1] = rng.integers(0, 1024, N)
        tr = rng.integers(0, 256, K).astype(np.int16)
        bl = rng.integers(0, 256, K).astype(np.int16)
        out[f"dc_top_{N}"], out[f"dc_left_{N}"] = top, left
        out[f"pl_tr_{N}"], out[f"pl_bl_{N}"] = tr, bl
        out[f"dc_pred_{N}"] = np.stack([R.intra_dc_predict(top[k], left[k], N) for k in range(K)])
        out[f"pl_pred_{N}"] = np.stack([
            R.intra_planar_predict(top[k], left[k], int(tr[k]), int(bl[k]), N) for k in range(K)])
        # angular: every mode, 2 random cases, corner argument != top[0] != left[0]
        A = 2
        atop = rng.integers(0, 256, (A, 2 * N + 1)).astype(np.int16)
        aleft = rng.integers(0, 256, (A, 2 * N + 1)).astype(np.int16)
        atop[1] = rng.integers(0, 1024, 2 * N + 1)
        aleft[1] = rng.integers(0, 1024, 2 * N + 1)
        acorner = rng.integers(0, 256, A).astype(np.int16)
        out[f"ang_top_{N}"], out[f"ang_left_{N}"], out[f"ang_corner_{N}"] = atop, aleft, acorner
        out[f"ang_pred_{N}"] = np.stack([
            np.stack([R.intra_angular_predict(atop[a], aleft[a], int(acorner[a]), m, N)
                      for m in range(2, 35)]) for a in range(A)])
        # short arrays (N+1 entries): replicate-last padding of the primary, skipped secondary
        stop = rng.integers(0, 256, N + 1).astype(np.int16)
        sleft = rng.integers(0, 256, N + 1).astype(np.int16)
        out[f"short_top_{N}"], out[f"short_left_{N}"] = stop, sleft
        out[f"short_pred_{N}"] = np.stack([
            R.intra_angular_predict(stop, sleft, int(stop[0]), m, N) for m in range(2, 35)])
    save("predictors.npz", **out)


# --------------------------------------------------------------- transforms
def g_transforms():
    rng = np.random.default_rng(99)
    out = {}
    for N in SIZES:
        for dst in ((False, True) if N == 4 else (False,)):
            tag = f"{N}{'dst' if dst else ''}"
            T = R.DST4 if dst else {4: R.DCT4, 8: R.DCT8, 16: R.DCT16, 32: R.DCT32}[N]
            cases = [
                rng.integers(-255, 256, (N, N)),
                rng.integers(-255, 256, (N, N)),
                rng.integers(-32768, 32768, (N, N)),           # full-range int16
                np.full((N, N), 255), np.full((N, N), -255),
                255 * np.sign(np.outer(T[N - 1], T[1])),        # basis pattern: max coefficient
                255 * ((np.indices((N, N)).sum(0) % 2) * 2 - 1),  # checkerboard
            ]
            imp = np.zeros((N, N), np.int64)
            imp[N // 2, 1] = 200
            cases.append(imp)
            x = np.stack(cases).astype(np.int16)
            fw = np.stack([R.forward_transform(c, use_dst=dst) for c in x])
            # inverse of (a) the forward output, (b) random small coefficients
            cin = np.concatenate([fw, rng.integers(-600, 601, (2, N, N)).astype(np.int32)])
            inv = np.stack([R.inverse_transform(c, use_dst=dst) for c in cin])
            out[f"x_{tag}"], out[f"fwd_{tag}"] = x, fw
            out[f"cin_{tag}"], out[f"inv_{tag}"] = cin.astype(np.int32), inv
    save("transforms.npz", **out)


# -------------------------------------------------------------------- quant
def g_quant():
    rng = np.random.default_rng(7)
    out = {}
    for N in SIZES:
        c = rng.integers(-1100, 1101, (N, N)).astype(np.int32)
        c.flat[:6] = [0, 1, -1, 32767, -32768, 170]
        big = rng.integers(-2**31, 2**31, (N, N)).astype(np.int32)
        big.flat[:2] = [-2**31, 2**31 - 1]
        lv = rng.integers(-300, 301, (N, N)).astype(np.int32)
        lv.flat[:5] = [0, 1, -1, 3, -3]
        out[f"c_{N}"], out[f"big_{N}"], out[f"lv_{N}"] = c, big, lv
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out[f"q_intra_{N}"] = np.stack([R.quantize(c, q, N, True) for q in range(-2, 54)])
            out[f"q_inter_{N}"] = np.stack([R.quantize(c, q, N, False) for q in range(-2, 54)])
            out[f"qbig_{N}"] = np.stack([R.quantize(big, q, N, True) for q in (0, 22, 51)])
            out[f"dq_{N}"] = np.stack([R.dequantize(lv, q, N) for q in range(-2, 54)])
            out[f"dqbig_{N}"] = np.stack([R.dequantize(big, q, N) for q in (0, 22, 51)])
    save("quant.npz", **out)


# ------------------------------------------------------------------ metrics
def g_metrics():
    rng = np.random.default_rng(11)
    out = {}
    a = rng.integers(0, 256, (8, 4, 4)).astype(np.int16)
    b = rng.integers(0, 256, (8, 4, 4)).astype(np.int16)
    out["a4"], out["b4"] = a, b
    out["sad4"] = np.array([RM.sad(a[i], b[i]) for i in range(8)], np.int64)
    out["satd4"] = np.array([RM.satd_4x4(a[i], b[i]) for i in range(8)], np.int64)
    out["energy4"] = np.array([RM.residual_energy(a[i] - b[i]) for i in range(8)], np.int64)
    A = rng.integers(0, 256, (48, 64)).astype(np.int16)
    B = np.clip(A + rng.integers(-9, 10, A.shape), 0, 255).astype(np.int16)
    out["A"], out["B"] = A, B
    out["sad_AB"] = np.array(RM.sad(A, B), np.int64)
    out["mse_AB"] = np.array(RM.mse(A, B), np.float64)
    out["psnr_AB"] = np.array(RM.psnr(A.astype(np.uint8), B.astype(np.uint8)), np.float64)
    out["psnr_same"] = np.array(RM.psnr(A, A), np.float64)
    save("metrics.npz", **out)


# ------------------------------------------- README quick-start (config 1, G1)
def g_readme():
    top = np.array([100, 102, 101, 99], np.int16)
    left = np.array([101, 100, 102, 103], np.int16)
    orig = np.array([[102, 101, 100, 100], [103, 102, 101, 100],
                     [103, 102, 100, 99], [104, 101, 99, 98]], np.int16)
    pred = R.intra_dc_predict(top, left, 4)
    res = R.residual_block(orig, pred)
    out = dict(top=top, left=left, orig=orig, pred=pred, res=res)
    for dst in (True, False):
        t = "dst" if dst else "dct"
        co = R.forward_transform(res, use_dst=dst)
        lv = R.quantize_block(co, 22)
        dq = R.dequantize_block(lv, 22)
        rr = R.inverse_transform(dq, use_dst=dst)
        rec = R.clip_to_pixel_range(R.reconstruct_block(pred, rr))
        out.update({f"coeff_{t}": co, f"levels_{t}": lv, f"dq_{t}": dq, f"rres_{t}": rr,
                    f"recon_{t}": rec})
    out["psnr"] = np.array(RM.psnr(orig, out["recon_dst"]), np.float64)
    out["sad"] = np.array(RM.sad(orig, pred), np.int64)
    out["satd"] = np.array(RM.satd_4x4(orig, pred), np.int64)
    save("readme.npz", **out)


# -------------------------------------------------- frame-level compositions
def synth_smooth(H, W, seed):
    """SURVEY 8d (iii) 'smooth': separable ramp + low-amplitude seeded noise."""
    rng = np.random.default_rng(4321 + seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = 40 + (150 * xx) // max(W - 1, 1) + (60 * yy) // max(H - 1, 1)
    return np.clip(base + rng.integers(-12, 13, (H, W)), 0, 255).astype(np.int16)


def ref_gather(plane, blk, N, T, L):
    """SURVEY 8a K1 gather oracle, literal."""
    bv = BlockView(plane=plane, x=blk.x, y=blk.y, size=N)
    c = bv.get_top_left_neighbor()
    top = np.concatenate([[c], bv.get_top_neighbors(T)]).astype(np.int16)
    left = np.concatenate([[c], bv.get_left_neighbors(L)]).astype(np.int16)
    return top, left, c


def pad(a, N):
    if len(a) < 2 * N + 1:
        a = np.concatenate([a, np.full(2 * N + 1 - len(a), a[-1], a.dtype)])
    return a


def ref_predict(top, left, c, mode, N):
    pt, pl = pad(top, N), pad(left, N)
    if mode == 1:
        return R.intra_dc_predict(top[1:N + 1], left[1:N + 1], N)
    if mode == 0:
        return R.intra_planar_predict(top[1:N + 1], left[1:N + 1], int(pt[N + 1]), int(pl[N + 1]), N)
    return R.intra_angular_predict(top, left, c, mode, N)


def ref_satd(a, b, N):
    return sum(RM.satd_4x4(a[y:y + 4, x:x + 4], b[y:y + 4, x:x + 4])
               for y in range(0, N, 4) for x in range(0, N, 4))


def ref_search(orig, top, left, c, N, cost):
    best = None
    costs = np.zeros(35, np.int32)
    for m in [1, 0] + list(range(2, 35)):
        p = ref_predict(top, left, c, m, N)
        k = RM.sad(orig, p) if cost == "sad" else ref_satd(orig, p, N)
        costs[m] = k
        if best is None or k < best[1]:
            best = (m, k, p)
    return best + (costs,)


def ref_pipeline(orig, pred, N, qp):
    dst = (N == 4)
    co = R.forward_transform(R.residual_block(orig, pred), use_dst=dst)
    lv = R.quantize_block(co, qp)
    rr = R.inverse_transform(R.dequantize_block(lv, qp), use_dst=dst)
    return co, lv, R.clip_to_pixel_range(R.reconstruct_block(pred, rr))


def ref_encode_frame(src, N, cost, qp, recon_neighbours):
    H, W = src.shape
    splane = Plane(data=src.copy())
    rplane = Plane.zeros(H, W, dtype=np.int16)
    rows = dict(modes=[], costs=[], pred=[], coeff=[], levels=[], recon=[], all_costs=[],
                top=[], left=[], corner=[])
    for blk in iterate_blocks(splane, N):
        if recon_neighbours:
            top, left, c = ref_gather(rplane, blk, N, 2 * N, N)
        else:
            top, left, c = ref_gather(splane, blk, N, 2 * N, 2 * N)
        orig = blk.copy_pixels()
        m, k, p, costs = ref_search(orig, top, left, c, N, cost)
        co, lv, rec = ref_pipeline(orig, p, N, qp)
        BlockView(plane=rplane, x=blk.x, y=blk.y, size=N).write_pixels(rec)
        for key, v in (("modes", m), ("costs", k), ("pred", p), ("coeff", co), ("levels", lv),
                       ("recon", rec), ("all_costs", costs), ("top", pad(top, N)),
                       ("left", pad(left, N)), ("corner", c)):
            rows[key].append(v)
    out = {k: np.array(v) for k, v in rows.items()}
    out["modes"] = out["modes"].astype(np.uint8)
    out["costs"] = out["costs"].astype(np.int32)
    out["corner"] = out["corner"].astype(np.int16)
    out["recon_plane"] = rplane.data
    out["psnr"] = np.array(RM.psnr(src.astype(np.uint8), rplane.data.astype(np.uint8)), np.float64)
    return out


def g_frames():
    out = {}
    # odd-shaped frame: partial blocks right and bottom, truncated neighbour slices
    for N, (H, W) in ((4, (22, 27)), (8, (22, 27)), (16, (40, 52)), (32, (70, 100))):
        src = synth_smooth(H, W, N)
        out[f"src_{N}"] = src
        for cost, qp in (("sad", 27), ("satd", 22)):
            for rn in (0, 1):
                if N == 32 and cost == "satd" and rn == 0:
                    continue
                r = ref_encode_frame(src, N, cost, qp, bool(rn))
                for k, v in r.items():
                    out[f"{k}_{N}_{cost}_{rn}"] = v
    # a noise frame (non-trivial levels) for N=4/8, SAD, both neighbour flavours
    rng = np.random.default_rng(1234)
    noise = rng.integers(0, 256, (24, 32), dtype=np.uint8).astype(np.int16)
    out["noise"] = noise
    for N in (4, 8):
        for rn in (0, 1):
            r = ref_encode_frame(noise, N, "sad", 22, bool(rn))
            for k, v in r.items():
                out[f"noise_{k}_{N}_{rn}"] = v
    save("frames.npz", **out)


def g_stats():
    """Level statistics (quant.py:153-178): estimate_bits / count_nonzero / is_all_zero."""
    from nano_hevc.quant import estimate_bits, count_nonzero, is_all_zero
    rng = np.random.default_rng(31)
    out = {}
    cases = [rng.integers(-300, 301, (8, 8)), rng.integers(-3, 4, (32, 32)), np.zeros((4, 4), np.int64),
             rng.integers(-13600, 13601, (16, 16)), (2 ** rng.integers(0, 12, (8, 8))) - 1,
             rng.integers(-1, 2, (1000, 64))]
    for i, c in enumerate(cases):
        c = c.astype(np.int32)
        out[f"lv_{i}"] = c
        out[f"bits_{i}"] = np.array(estimate_bits(c), np.int64)
        out[f"nnz_{i}"] = np.array(count_nonzero(c), np.int64)
        out[f"zero_{i}"] = np.array(bool(is_all_zero(c)))
    out["n_cases"] = np.array(len(cases), np.int64)
    save("stats.npz", **out)


def g_cli():
    """Whole-program goldens: encode_frame_intra (__main__.py:142-189) on create_test_frame."""
    out = {}
    for (H, W, bs) in ((64, 64, 8), (96, 128, 16)):
        fr = create_test_frame(H, W)
        rec, stats = encode_frame_intra(fr, bs)
        tag = f"{H}x{W}_{bs}"
        out[f"y_{tag}"] = fr.y.data
        out[f"recon_y_{tag}"] = rec.y.data
        out[f"stats_{tag}"] = np.array([stats["dc"], stats["planar"], stats["blocks"]], np.int64)
        out[f"psnr_y_{tag}"] = np.array(
            RM.psnr(fr.y.data.astype(np.uint8), rec.y.data.astype(np.uint8)), np.float64)
    save("cli.npz", **out)


def g_containers():
    """Frame containers (frame.py:121-308): PackedFrame.from_yuv420p / to_yuv420p (astype(np.uint8)
    wrap-around) and the FrameBufferPool acquire / release sequence."""
    from nano_hevc.frame import PackedFrame, FrameBufferPool, Frame
    out = {}
    rng = np.random.default_rng(2024)
    for (H, W) in ((6, 10), (18, 34), (64, 96)):
        raw = rng.integers(0, 256, H * W + 2 * (H // 2) * (W // 2), dtype=np.uint8).tobytes()
        pf = PackedFrame.from_yuv420p(raw, H, W)
        tag = f"{H}x{W}"
        out[f"raw_{tag}"] = np.frombuffer(raw, np.uint8)
        out[f"y_{tag}"], out[f"u_{tag}"], out[f"v_{tag}"] = pf.y.copy(), pf.u.copy(), pf.v.copy()
        assert pf.to_yuv420p() == raw and Frame.from_yuv420p(raw, H, W).to_yuv420p() == raw
        # an int16 frame with samples outside 0..255: to_yuv420p keeps the low 8 bits
        p16 = PackedFrame(H, W, dtype=np.int16)
        vals = rng.integers(-32768, 32768, p16._buffer.size).astype(np.int16)
        vals[:8] = [0, 255, 256, -1, -256, 300, 32767, -32768]
        np.copyto(p16._buffer, vals)
        out[f"i16_{tag}"] = vals
        out[f"i16_bytes_{tag}"] = np.frombuffer(p16.to_yuv420p(), np.uint8)
    # pool bookkeeping: a scripted sequence of operations and what the reference answers
    pool = FrameBufferPool(8, 8, pool_size=3)
    trace = []
    def snap(tag, val):
        trace.append((tag, val, pool.available_count, pool.in_use_count))
    a, _ = pool.acquire(); snap("acquire", a)
    b, fb = pool.acquire(); snap("acquire", b)
    fb.y[:] = 7
    pool.release(a); snap("release", a)
    c, _ = pool.acquire(); snap("acquire", c)
    d, _ = pool.acquire(); snap("acquire", d)
    try:
        pool.acquire()
    except RuntimeError as e:
        out["pool_exhausted_msg"] = np.array(str(e))
    try:
        pool.release(a + 100)
    except ValueError as e:
        out["pool_release_msg"] = np.array(str(e))
    pool.release(b); snap("release", b)
    e2, fe = pool.acquire(clear=False); snap("acquire", e2)
    out["pool_kept_value"] = np.array(int(fe.y[0, 0]))      # clear=False keeps the 7 written above
    e3 = pool.release(e2); snap("release", e2)
    _, fz = pool.acquire(clear=True)
    out["pool_cleared_value"] = np.array(int(fz.y[0, 0]))
    out["pool_trace"] = np.array([[0 if t == "acquire" else 1, v, av, iu] for (t, v, av, iu) in trace], np.int64)
    save("containers.npz", **out)


if __name__ == "__main__":
    g_tables()
    g_predictors()
    g_transforms()
    g_quant()
    g_metrics()
    g_readme()
    g_cli()
    g_stats()
    g_frames()
    g_containers()
