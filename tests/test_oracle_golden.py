"""Pin the C oracle (oracle/nh_oracle.c) to the reference.

Golden arrays come from tests/golden/make_golden.py, which imports the
unmodified reference (nano_hevc.*) in the authoring container.  Known-answer
values are those of the reference's own tests.
"""
import numpy as np
import pytest

import oracle as O
from conftest import golden

SIZES = (4, 8, 16, 32)


def test_tables_match_reference():
    g = golden("tables.npz")
    for n in SIZES:
        assert np.array_equal(O.get_matrix(n), g[f"DCT{n}"])
    assert np.array_equal(O.get_matrix(4, True), g["DST4"])
    assert [O.intra_pred_angle(m) for m in range(2, 35)] == list(g["INTRA_PRED_ANGLE"])
    # tests/test_intra_angular.py:190-197 spot values
    assert O.intra_pred_angle(2) == 32 and O.intra_pred_angle(10) == 0
    assert O.intra_pred_angle(18) == -32 and O.intra_pred_angle(26) == 0 and O.intra_pred_angle(34) == 32
    qp = [O.get_qp_params(q) for q in range(-3, 56)]
    assert np.array_equal(np.array(qp), g["qp_params"])  # includes clamping (tests/test_quant.py:28-56)
    with pytest.raises(ValueError):
        O.get_matrix(12)


@pytest.mark.parametrize("n", SIZES)
def test_predictors(n):
    g = golden("predictors.npz")
    top, left = g[f"dc_top_{n}"], g[f"dc_left_{n}"]
    for k in range(top.shape[0]):
        assert np.array_equal(O.intra_dc_predict(top[k], left[k], n), g[f"dc_pred_{n}"][k])
        assert np.array_equal(
            O.intra_planar_predict(top[k], left[k], g[f"pl_tr_{n}"][k], g[f"pl_bl_{n}"][k], n),
            g[f"pl_pred_{n}"][k])
    at, al, ac = g[f"ang_top_{n}"], g[f"ang_left_{n}"], g[f"ang_corner_{n}"]
    for a in range(at.shape[0]):
        for m in range(2, 35):
            got = O.intra_angular_predict(at[a], al[a], ac[a], m, n)
            assert got.dtype == np.int16
            assert np.array_equal(got, g[f"ang_pred_{n}"][a, m - 2]), (n, a, m)
    st, sl = g[f"short_top_{n}"], g[f"short_left_{n}"]
    for m in range(2, 35):
        assert np.array_equal(O.intra_angular_predict(st, sl, st[0], m, n),
                              g[f"short_pred_{n}"][m - 2]), (n, m)


def test_reference_known_answers_intra():
    # tests/test_intra_dc.py:23-56
    top = np.array([100, 102, 101, 99], np.int16)
    left = np.array([101, 100, 102, 103], np.int16)
    assert np.all(O.intra_dc_predict(top, left, 4) == 101)
    # tests/test_intra_planar.py:56-76 corners
    top = np.array([0, 0, 0, 0], np.int16)
    left = np.array([0, 0, 0, 0], np.int16)
    p = O.intra_planar_predict(top, left, 255, 255, 4)
    assert p.shape == (4, 4)
    # tests/test_intra_angular.py:69-85 -- the non-standard negative-angle projection (Q3)
    top = np.array([0, 10, 20, 30, 40, 50, 60, 70, 80], np.int16)
    left = np.array([0, 5, 5, 5, 5, 5, 5, 5, 5], np.int16)
    exp = np.array([[0, 10, 20, 30], [0, 0, 10, 20], [5, 0, 0, 10], [5, 5, 0, 0]], np.int16)
    assert np.array_equal(O.intra_angular_predict(top, left, 0, 18, 4), exp)
    # tests/test_intra_angular.py:25-43 mode 26 with 9-entry arrays at size 8 (replicate-last)
    top = np.array([99, 100, 110, 120, 130, 0, 0, 0, 0], np.int16)
    left = np.array([99, 50, 50, 50, 50, 0, 0, 0, 0], np.int16)
    for n in (4, 8):
        p = O.intra_angular_predict(top, left, 99, 26, n)
        assert [int(v) for v in p[0, :4]] == [100, 110, 120, 130] and np.all(p[:, 0] == 100)
    # tests/test_intra_angular.py:45-67 and :111-133
    top = np.array([0, 10, 20, 30, 40, 50, 60, 70, 80], np.int16)
    p = O.intra_angular_predict(top, np.zeros(9, np.int16), 0, 34, 4)
    assert (p[0, 0], p[0, 3], p[1, 0], p[3, 3]) == (20, 50, 30, 80)
    p = O.intra_angular_predict(np.zeros(9, np.int16), top, 0, 2, 4)
    assert (p[0, 0], p[3, 0], p[0, 1], p[3, 3]) == (20, 50, 30, 80)
    # all 33 modes: uniform refs -> uniform prediction (tests/test_intra_angular.py:175-188)
    u = np.full(9, 128, np.int16)
    for m in range(2, 35):
        assert np.all(O.intra_angular_predict(u, u, 128, m, 4) == 128)


@pytest.mark.parametrize("tag", ["4", "4dst", "8", "16", "32"])
def test_transforms(tag):
    g = golden("transforms.npz")
    dst = tag.endswith("dst")
    for x, f in zip(g[f"x_{tag}"], g[f"fwd_{tag}"]):
        got = O.forward_transform(x, dst)
        assert got.dtype == np.int32 and np.array_equal(got, f)
    for c, r in zip(g[f"cin_{tag}"], g[f"inv_{tag}"]):
        assert np.array_equal(O.inverse_transform(c, dst), r)


def test_transform_bad_size():
    with pytest.raises(ValueError):
        O.forward_transform(np.zeros((5, 5), np.int16))
    with pytest.raises(ValueError):
        O.inverse_transform(np.zeros((64, 64), np.int32))


@pytest.mark.parametrize("n", SIZES)
def test_quant(n):
    g = golden("quant.npz")
    c, big, lv = g[f"c_{n}"], g[f"big_{n}"], g[f"lv_{n}"]
    for i, qp in enumerate(range(-2, 54)):
        assert np.array_equal(O.quantize(c, qp, n, True), g[f"q_intra_{n}"][i]), qp
        assert np.array_equal(O.quantize(c, qp, n, False), g[f"q_inter_{n}"][i]), qp
        assert np.array_equal(O.dequantize(lv, qp, n), g[f"dq_{n}"][i]), qp
    for i, qp in enumerate((0, 22, 51)):
        assert np.array_equal(O.quantize(big, qp, n, True), g[f"qbig_{n}"][i])
        assert np.array_equal(O.dequantize(big, qp, n), g[f"dqbig_{n}"][i])


def test_quant_known_answers():
    # tests/test_quant.py:70-77: 5 @ QP40 -> 0
    assert np.all(O.quantize(np.full((4, 4), 5, np.int32), 40, 4) == 0)
    # SURVEY G2 dequant asymmetry (arithmetic shift on negatives)
    for qp, lvl, exp in ((0, 1, 3), (0, -1, -2), (0, 3, 8), (0, -3, -7), (5, 1, 5), (5, -1, -4),
                         (22, 1, 32), (22, -1, -32), (51, 1, 912), (51, -1, -912)):
        assert int(O.dequantize(np.array([[lvl]], np.int32), qp)[0, 0]) == exp
    # first surviving |c| at QP22
    for n, first in ((4, 22), (8, 43), (16, 86), (32, 171)):
        c = np.array([[first - 1, first]], np.int32)
        assert list(O.quantize(c, 22, n)[0]) == [0, 1]


def test_metrics():
    g = golden("metrics.npz")
    for i in range(8):
        assert O.sad(g["a4"][i], g["b4"][i]) == g["sad4"][i]
        assert O.satd_4x4(g["a4"][i], g["b4"][i]) == g["satd4"][i]
        assert O.residual_energy(g["a4"][i] - g["b4"][i]) == g["energy4"][i]
    assert O.sad(g["A"], g["B"]) == int(g["sad_AB"])
    assert O.mse(g["A"], g["B"]) == float(g["mse_AB"])
    assert O.psnr(g["A"], g["B"]) == float(g["psnr_AB"])
    assert O.psnr(g["A"], g["A"]) == float("inf")


def test_readme_quickstart():
    g = golden("readme.npz")
    pred = O.intra_dc_predict(g["top"], g["left"], 4)
    assert np.array_equal(pred, g["pred"]) and np.all(pred == 101)
    res = O.residual_block(g["orig"], pred)
    assert np.array_equal(res, g["res"])
    # SURVEY G1
    assert np.array_equal(O.forward_transform(res, True),
                          [[-2, 5, 2, 1], [2, 0, 0, 0], [-2, 0, 0, 0], [-1, 0, 1, -1]])
    for t, dst in (("dst", True), ("dct", False)):
        co, lv, rec = O.block_pipeline(g["orig"], pred, 22, True, dst)
        assert np.array_equal(co, g[f"coeff_{t}"])
        assert np.array_equal(lv, g[f"levels_{t}"])
        assert np.array_equal(rec, g[f"recon_{t}"])
    assert O.psnr(g["orig"], g["recon_dst"]) == pytest.approx(44.044164868040994, rel=1e-12)
    assert O.sad(g["orig"], pred) == 21 and O.satd_4x4(g["orig"], pred) == 66


FRAME_CASES = [(n, c, rn) for n in SIZES for c in ("sad", "satd") for rn in (0, 1)
               if not (n == 32 and c == "satd" and rn == 0)]


@pytest.mark.parametrize("n,cost,rn", FRAME_CASES)
def test_frame_coder(n, cost, rn):
    g = golden("frames.npz")
    src = g[f"src_{n}"]
    qp = 27 if cost == "sad" else 22
    out = O.encode_frame(src, n, cost, qp, bool(rn), threads=2)
    for k in ("modes", "costs", "pred", "coeff", "levels", "recon", "recon_plane"):
        assert np.array_equal(out[k], g[f"{k}_{n}_{cost}_{rn}"]), (k, n, cost, rn)
    assert O.psnr(src, out["recon_plane"]) == pytest.approx(float(g[f"psnr_{n}_{cost}_{rn}"]), rel=1e-12)
    # K1 gather + per-mode costs
    T, L = 2 * n, (n if rn else 2 * n)
    plane = g[f"recon_plane_{n}_{cost}_{rn}"] if rn else src
    if not rn:
        top, left, corner = O.gather_refs_frame(plane, n, T, L)
        assert np.array_equal(top, g[f"top_{n}_{cost}_{rn}"])
        assert np.array_equal(left, g[f"left_{n}_{cost}_{rn}"])
        assert np.array_equal(corner, g[f"corner_{n}_{cost}_{rn}"])
        blocks = O.blocks_from_plane(src, n)
        for b in range(blocks.shape[0]):
            m, c, _, costs = O.search_block(blocks[b], top[b], left[b], corner[b], cost)
            assert m == g[f"modes_{n}_{cost}_{rn}"][b] and c == g[f"costs_{n}_{cost}_{rn}"][b]
            assert np.array_equal(costs, g[f"all_costs_{n}_{cost}_{rn}"][b])


@pytest.mark.parametrize("n,rn", [(4, 0), (4, 1), (8, 0), (8, 1)])
def test_frame_coder_noise(n, rn):
    g = golden("frames.npz")
    out = O.encode_frame(g["noise"], n, "sad", 22, bool(rn))
    for k in ("modes", "costs", "pred", "coeff", "levels", "recon", "recon_plane"):
        assert np.array_equal(out[k], g[f"noise_{k}_{n}_{rn}"]), (k, n, rn)
    assert np.count_nonzero(out["levels"]) > 0


def test_batched_matches_per_block():
    rng = np.random.default_rng(5)
    for n in SIZES:
        B = 9
        orig = rng.integers(0, 256, (B, n, n)).astype(np.int16)
        top = rng.integers(0, 256, (B, n)).astype(np.int16)
        left = rng.integers(0, 256, (B, n)).astype(np.int16)
        tr = rng.integers(0, 256, B).astype(np.int16)
        bl = rng.integers(0, 256, B).astype(np.int16)
        modes = rng.integers(0, 2, B).astype(np.uint8)
        p, c, l, r = O.pipeline_dcplanar_batch(orig, top, left, tr, bl, modes, 22, threads=3)
        for b in range(B):
            pb = (O.intra_dc_predict(top[b], left[b], n) if modes[b] == 1 else
                  O.intra_planar_predict(top[b], left[b], tr[b], bl[b], n))
            cb, lb, rb = O.block_pipeline(orig[b], pb, 22)
            assert np.array_equal(p[b], pb) and np.array_equal(c[b], cb)
            assert np.array_equal(l[b], lb) and np.array_equal(r[b], rb)


def test_level_statistics():
    g = golden("stats.npz")
    for i in range(int(g["n_cases"])):
        lv = g[f"lv_{i}"]
        assert O.estimate_bits(lv) == int(g[f"bits_{i}"])
        assert O.count_nonzero(lv) == int(g[f"nnz_{i}"])
        assert (O.count_nonzero(lv) == 0) == bool(g[f"zero_{i}"])
