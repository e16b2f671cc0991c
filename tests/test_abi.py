"""The C-ABI boundary without a GPU: libnh_b200.so loads, exports every symbol include/nh_b200.h
declares, the ctypes prototypes cover them all, host-side tables match the reference's, and compute
entry points fail loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nh_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nano_hevc_b200 import _lib
    handle = C.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/nh_b200.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes and header disagree"
    assert _lib.lib().nh_version() >= 100


def test_no_oracle_in_product_path():
    """The product package must never import / link the CPU oracle."""
    pkg = os.path.join(ROOT, "nano_hevc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "nh_oracle" not in text and "nho_" not in text, f


def test_host_tables_match_reference():
    import nano_hevc_b200 as P
    g = golden("tables.npz")
    for name in ("DCT4", "DCT8", "DCT16", "DCT32", "DST4"):
        got = getattr(P, name)
        assert got.dtype == np.int32 and np.array_equal(got, g[name]), name
    assert P.INTRA_PRED_ANGLE == list(g["INTRA_PRED_ANGLE"])
    assert P.QUANT_SCALE == list(g["QUANT_SCALE"]) and P.DEQUANT_SCALE == list(g["DEQUANT_SCALE"])
    assert [P.get_qp_params(q) for q in range(-3, 56)] == [tuple(r) for r in g["qp_params"]]


def test_error_codes_without_device():
    from nano_hevc_b200 import _lib
    L = _lib.lib()
    out = np.zeros(144, np.int32)
    assert L.nh_get_transform_matrix(12, 0, out.ctypes.data_as(C.c_void_p)) == _lib.NH_E_SIZE
    assert "Unsupported transform size: 12" in _lib.last_error()
    a = C.c_int()
    assert L.nh_get_intra_pred_angle(35, C.byref(a)) == _lib.NH_E_ARG
    assert L.nh_forward_transform(None, 0, None, 1, 8, 0, None) == _lib.NH_E_ARG
    assert L.nh_forward_transform(None, 0, None, 1, 7, 0, None) == _lib.NH_E_SIZE
    with pytest.raises(ValueError, match="Unsupported transform size"):
        _lib.check(_lib.NH_E_SIZE)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_kernel_selection_and_accounting_entry_points():
    """Entry points without a reference counterpart: argument validation works without a device."""
    import ctypes as C
    from nano_hevc_b200 import _lib
    L = _lib.lib()
    for good in (1, 2):
        assert L.nh_set_rows_impl(good) == 0
    for bad in (0, 3, -1):
        assert L.nh_set_rows_impl(bad) != 0 and b"nh_set_rows_impl" in L.nh_last_error()
    L.nh_set_rows_impl(2)
    for good in (1, 2, 3, 4, 5, 6):
        assert L.nh_set_search_impl(good) == 0
    assert L.nh_set_search_impl(7) != 0 and b"nh_set_search_impl" in L.nh_last_error()
    L.nh_set_search_impl(2)
    for w, b in ((0, 0), (1, 2), (2, 0), (4, 1), (8, 3), (12, 0)):
        assert L.nh_set_wave_impl(w, b) == 0
    for w, b in ((3, 0), (16, 0), (-1, 0), (0, 4), (0, -1)):
        assert L.nh_set_wave_impl(w, b) != 0 and b"nh_set_wave_impl" in L.nh_last_error()
    L.nh_set_wave_impl(0, 0)
    for good in (1, 2, 3, 4):
        assert L.nh_set_fused_impl(good) == 0
    assert L.nh_set_fused_impl(5) != 0
    L.nh_set_fused_impl(4)
    up, down = C.c_int64(-1), C.c_int64(-1)
    assert L.nh_host_pipeline_last_transfer(C.byref(up), C.byref(down)) == 0 and up.value >= 0 and down.value >= 0
    assert L.nh_host_pipeline_last_transfer(None, None) == 0
    assert L.nh_host_pipeline_scratch_bytes(8, 1024) > 0 and L.nh_host_pipeline_scratch_bytes(12, 1024) == 0
    assert L.nh_convert_u8_to_i16(None, None, 0, None) == 0       # empty is fine
    assert L.nh_convert_u8_to_i16(None, None, 5, None) != 0       # null pointers are not
    assert L.nh_convert_i16_to_u8(None, None, -1, None) != 0


def test_compute_fails_loudly_without_gpu():
    import nano_hevc_b200 as P
    from nano_hevc_b200 import batched
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.forward_transform(np.zeros((4, 4), np.int16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        batched.forward_transform_batched(torch.zeros((1, 4, 4), dtype=torch.int16))
    assert P._lib.lib().nh_device_ok() == 0


def test_angular_mode_wraparound_rule():
    """intra.py:142-143: modes < 2 index INTRA_PRED_ANGLE from the end and run horizontally."""
    import nano_hevc_b200 as P
    ang = P.INTRA_PRED_ANGLE
    for mode in range(-31, 2):
        if mode == -15:
            with pytest.raises(ValueError):
                P._equivalent_angular_mode(mode)
            continue
        eq = P._equivalent_angular_mode(mode)
        assert 2 <= eq < 18 and ang[eq - 2] == ang[mode - 2]
    for bad in (35, 40, -32):
        with pytest.raises(IndexError):
            P._equivalent_angular_mode(bad)
