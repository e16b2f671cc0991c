"""Replay of the reference's own test-suite against the per-block API.

tests/golden/ref_test_calls.npz holds every distinct hot-path call that the unmodified suite of the
reference (tests/test_intra_dc.py, test_intra_planar.py, test_intra_angular.py, test_transform.py,
test_quant.py -- 77 tests) makes, with the value the reference returned, recorded by
tests/golden/record_reference_tests.py (the test sources themselves are not copied and the reference
is not present on the GPU box).  Each call is replayed
  * through nano_hevc_b200's per-block numpy API on the GPU (same names, positional / keyword
    parameters) and
  * through the C oracle on the CPU,
and must give the same value, shape and dtype.  Since the reference's suite passes on the reference,
every exact pin in it is implied by this equality."""
import json

import numpy as np
import pytest

from conftest import golden


def _calls():
    z = golden("ref_test_calls.npz")
    index = json.loads(bytes(z["index"]).decode())

    def get(x):
        return z[x["array"]] if "array" in x else x["scalar"]

    out = []
    for i, e in enumerate(index):
        args = [get(a) for a in e["args"]]
        kwargs = {k: get(v) for k, v in e["kwargs"].items()}
        want = e["value"] if e["kind"] == "raises" else get(e["value"])
        out.append(pytest.param(e["fn"], args, kwargs, e["kind"], want, id=f"{i}-{e['fn']}"))
    return out


CALLS = _calls()


def _same(got, want, fn):
    if isinstance(want, np.ndarray):
        assert isinstance(got, np.ndarray), (fn, type(got))
        assert got.dtype == want.dtype, (fn, got.dtype, want.dtype)
        assert got.shape == want.shape, (fn, got.shape, want.shape)
        assert np.array_equal(got, want), (fn, got, want)
    elif isinstance(want, list):   # get_qp_params -> (per, rem)
        assert list(got) == want, (fn, got, want)
    elif isinstance(want, float):
        assert got == pytest.approx(want, rel=1e-12), (fn, got, want)
    else:
        assert type(got) is type(want) and got == want, (fn, got, want)


def test_recording_covers_the_suite():
    fns = {p.values[0] for p in CALLS}
    assert len(CALLS) >= 160
    assert {"intra_dc_predict_4x4", "intra_dc_predict", "intra_planar_predict", "intra_angular_predict",
            "residual_block", "reconstruct_block", "clip_to_pixel_range", "forward_transform",
            "inverse_transform", "quantize", "dequantize", "get_qp_params"} <= fns


@pytest.mark.parametrize("fn,args,kwargs,kind,want", CALLS)
def test_oracle_replays_reference_suite(fn, args, kwargs, kind, want):
    import oracle as O
    alias = {"intra_dc_predict_4x4": lambda top, left: O.intra_dc_predict(top, left, 4),
             "is_all_zero": lambda lv: O.count_nonzero(lv) == 0}
    for n in (4, 8, 16, 32):
        alias[f"forward_transform_{n}x{n}"] = O.forward_transform
        alias[f"inverse_transform_{n}x{n}"] = O.inverse_transform
    f = alias.get(fn) or getattr(O, fn)
    if kind == "raises":
        with pytest.raises(Exception) as ei:
            f(*args, **kwargs)
        assert type(ei.value).__name__ == want
        return
    _same(f(*args, **kwargs), want, fn)


@pytest.mark.gpu
@pytest.mark.parametrize("fn,args,kwargs,kind,want", CALLS)
def test_gpu_api_replays_reference_suite(fn, args, kwargs, kind, want):
    import nano_hevc_b200 as P
    f = getattr(P, fn)
    if kind == "raises":
        with pytest.raises(Exception) as ei:
            f(*args, **kwargs)
        assert type(ei.value).__name__ == want
        return
    _same(f(*args, **kwargs), want, fn)
