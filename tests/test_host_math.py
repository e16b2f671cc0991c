"""The integer code the kernels use (nano_hevc_b200/csrc/nh_math.cuh), compiled for the HOST by
tests/shim/host_shim.cpp, against the golden vectors of the reference and the C oracle.  This runs
without a GPU and catches arithmetic mistakes before any GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from conftest import ROOT, golden

SIZES = (4, 8, 16, 32)


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("shim") / "libhost_shim.so"
    src = os.path.join(ROOT, "tests", "shim", "host_shim.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fwrapv", "-o", str(out), src])
    return C.CDLL(str(out))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _xf(shim, x, n, dst, inv):
    x = np.ascontiguousarray(x, np.int32)
    out = np.empty_like(x)
    assert shim.shim_transform2d(n, int(dst), int(inv), _p(x), _p(out)) == 0
    return out


def test_tables(shim):
    g = golden("tables.npz")
    for n in SIZES:
        m = np.array([[shim.shim_matrix(n, i, j) for j in range(n)] for i in range(n)])
        assert np.array_equal(m, g[f"DCT{n}"])
    assert [shim.shim_angle(m) for m in range(2, 35)] == list(g["INTRA_PRED_ANGLE"])
    for a, inv in ((-2, -4096), (-5, -1638), (-9, -910), (-13, -630), (-17, -482), (-21, -390),
                   (-26, -315), (-32, -256), (0, 0), (5, 0)):
        assert shim.shim_inv_angle(a) == inv


@pytest.mark.parametrize("tag", ["4", "4dst", "8", "16", "32"])
def test_butterfly_transforms_match_reference(shim, tag):
    g = golden("transforms.npz")
    n, dst = int(tag.replace("dst", "")), tag.endswith("dst")
    for x, want in zip(g[f"x_{tag}"], g[f"fwd_{tag}"]):
        assert np.array_equal(_xf(shim, x, n, dst, False), want)
    for c, want in zip(g[f"cin_{tag}"], g[f"inv_{tag}"]):
        assert np.array_equal(_xf(shim, c, n, dst, True), want)
    # full-range int32 inputs: wrap-around accumulators must agree with the oracle's loop form
    rng = np.random.default_rng(5)
    for _ in range(4):
        x = rng.integers(-2**31, 2**31, (n, n)).astype(np.int32)
        assert np.array_equal(_xf(shim, x, n, dst, False), O.forward_transform(x, dst))
        assert np.array_equal(_xf(shim, x, n, dst, True), O.inverse_transform(x, dst))


@pytest.mark.parametrize("n", SIZES)
def test_quant_dequant(shim, n):
    g = golden("quant.npz")
    l2 = {4: 2, 8: 3, 16: 4, 32: 5}[n]
    c, lv, big = g[f"c_{n}"], g[f"lv_{n}"], g[f"big_{n}"]
    out = np.empty_like(c)
    for i, qp in enumerate(range(-2, 54)):
        shim.shim_quant(_p(c), _p(out), C.c_int64(c.size), qp, l2, 1)
        assert np.array_equal(out, g[f"q_intra_{n}"][i]), qp
        shim.shim_quant(_p(c), _p(out), C.c_int64(c.size), qp, l2, 0)
        assert np.array_equal(out, g[f"q_inter_{n}"][i]), qp
        shim.shim_dequant(_p(lv), _p(out), C.c_int64(lv.size), qp)
        assert np.array_equal(out, g[f"dq_{n}"][i]), qp
    for i, qp in enumerate((0, 22, 51)):
        shim.shim_quant(_p(big), _p(out), C.c_int64(big.size), qp, l2, 1)
        assert np.array_equal(out, g[f"qbig_{n}"][i]), qp
        shim.shim_dequant(_p(big), _p(out), C.c_int64(big.size), qp)
        assert np.array_equal(out, g[f"dqbig_{n}"][i]), qp


@pytest.mark.parametrize("n", SIZES)
def test_predictors(shim, n):
    g = golden("predictors.npz")
    at, al, ac = g[f"ang_top_{n}"], g[f"ang_left_{n}"], g[f"ang_corner_{n}"]
    out = np.empty((n, n), np.int16)
    for a in range(at.shape[0]):
        for m in range(2, 35):
            shim.shim_predict_mode(n, _p(at[a]), _p(al[a]), int(ac[a]), m, _p(out))
            assert np.array_equal(out, g[f"ang_pred_{n}"][a, m - 2]), (n, a, m)
        for m in (0, 1):
            shim.shim_predict_mode(n, _p(at[a]), _p(al[a]), int(ac[a]), m, _p(out))
            assert np.array_equal(out, O.predict_mode(at[a], al[a], ac[a], m, n))


@pytest.mark.parametrize("n", SIZES)
def test_fast_quant_is_exact_in_the_pixel_domain(shim, n):
    """32-bit quant / dequant (kernel fast path) == the int64 oracle for every coefficient the
    forward transform can produce from residuals in [-4095, 4095] and every QP."""
    l2 = {4: 2, 8: 3, 16: 4, 32: 5}[n]
    rng = np.random.default_rng(n)
    c = np.concatenate([np.arange(-32394, 32395, 7), rng.integers(-32394, 32395, 4000),
                        [-32394, 32394, -1, 0, 1]]).astype(np.int32)
    out = np.empty_like(c)
    for qp in range(0, 52):
        for intra in (1, 0):
            shim.shim_quant_fast(_p(c), _p(out), C.c_int64(c.size), qp, l2, intra)
            want = O.quantize(c, qp, n, bool(intra))
            assert np.array_equal(out, want), (qp, intra)
        lv = O.quantize(c, qp, n, True)
        shim.shim_dequant_fast(_p(lv), _p(out), C.c_int64(lv.size), qp)
        assert np.array_equal(out, O.dequantize(lv, qp)), qp


def _xf_dp(shim, x, n, inv):
    x = np.ascontiguousarray(x, np.int32)
    out = np.empty_like(x)
    assert shim.shim_transform2d_dp(n, int(inv), _p(x), _p(out)) == 0
    return out


@pytest.mark.parametrize("n", (16, 32))
def test_dp2a_butterflies_exact_in_the_pixel_domain(shim, n):
    """The IDP.2A butterflies (int16-lane operands, emulated WITH lane truncation on the host) must
    equal the reference through the whole fused chain for residuals in [-4095, 4095]: worst-case
    sign patterns of every basis function pair, random blocks, and every QP."""
    rng = np.random.default_rng(n)
    T = O.get_matrix(n)
    cases = [rng.integers(-4095, 4096, (n, n)) for _ in range(6)]
    for i in (0, 1, 2, 3, n // 2, n - 3, n - 2, n - 1):
        for j in (0, 1, n // 2 - 1, n - 2, n - 1):
            cases.append(4095 * np.sign(np.outer(T[i], T[j]) + 0.5).astype(np.int64))   # max |coeff[i][j]|
            cases.append(-4095 * np.sign(np.outer(T[i], T[j]) + 0.5).astype(np.int64))
    cases.append(4095 * ((np.indices((n, n)).sum(0) % 2) * 2 - 1))
    for res in cases:
        res = res.astype(np.int32)
        want = O.forward_transform(res)
        assert np.array_equal(_xf_dp(shim, res, n, False), want)
        assert np.abs(want).max() <= 32394
        for qp in (0, 5, 17, 23, 24, 30, 37, 45, 51):
            dq = O.dequantize(O.quantize(want, qp, n), qp)
            assert np.array_equal(_xf_dp(shim, dq, n, True), O.inverse_transform(dq)), qp
    # the emulation really truncates: an operand beyond int16 must break the equality
    big = np.zeros((n, n), np.int32)
    big[0, 0], big[n - 1, 0] = 30000, -30000
    assert not np.array_equal(_xf_dp(shim, big, n, False), O.forward_transform(big))


@pytest.mark.parametrize("n", (8, 16, 32))
def test_mma_operand_bounds(n):
    """The tensor-core kernels (csrc/nh_fused_mma.cuh) feed the four transform passes to f16 x f16 ->
    f32 MMAs.  That is exact only while every operand is an integer of magnitude <= 2048 and every
    accumulator stays below 2^24; the biased operand form additionally needs the value in
    [-512, 511].  Recompute the worst cases for 8-bit samples from the reference's own tables and
    quantiser (all QPs, intra and inter) and pin the numbers the kernels rely on."""
    g = golden("tables.npz")
    T = g[f"DCT{n}"].astype(np.int64)
    sh = int(np.log2(n)) + 5
    row_l1, col_l1 = int(np.abs(T).sum(1).max()), int(np.abs(T).sum(0).max())
    # rows other than the DC row sum to zero: the forward second pass's bias correction is 1536 * 64N on v = 0 only
    assert T.sum(1)[0] == 64 * n and not T.sum(1)[1:].any()
    res_in = 255
    temp = (row_l1 * res_in + (1 << (sh - 1))) >> sh           # first forward pass
    coeff = (row_l1 * temp + (1 << (sh - 1))) >> sh             # second forward pass
    c = np.arange(-coeff, coeff + 1, dtype=np.int32)
    dq = 0
    for qp in range(52):
        for intra in (True, False):
            dq = max(dq, int(np.abs(O.dequantize(O.quantize(c, qp, n, intra), qp, n)).max()))
    tmp2 = (col_l1 * dq + (1 << (sh - 1))) >> sh                # first inverse pass
    res = (col_l1 * tmp2 + (1 << (sh - 1))) >> sh               # second inverse pass
    bias = 1536
    acc_max = max(row_l1 * res_in, row_l1 * (temp + bias) + bias * 64 * n, col_l1 * (dq + bias) + bias * col_l1,
                  col_l1 * (tmp2 + bias) + bias * col_l1) + (1 << (sh - 1))
    assert acc_max < (1 << 24), acc_max
    assert max(temp, dq, tmp2) + bias <= 2048 + bias and temp <= 511  # temp always travels biased
    want = {8: (511, 1023, 720, 1348, 2523), 16: (511, 1023, 360, 661, 1214), 32: (511, 1023, 180, 328, 597)}[n]
    got = (temp, coeff, dq, tmp2, res)
    assert all(a <= b for a, b in zip(got, want)), (got, want)
    # which operands may use the biased form (value + 512 must fit 10 bits), as coded in the kernels
    assert (dq <= 511) == (n >= 16) and (tmp2 <= 511) == (n == 32)
    # plain-form operands must still be exact f16 integers; reconstruction bias keeps res + bias > 0
    assert max(dq, tmp2) <= 2048 and res < (4096 if n == 8 else 2048)
