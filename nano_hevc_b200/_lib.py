"""ctypes binding of ``libnh_b200.so`` (the sm_100a kernel library, ``include/nh_b200.h``).

There is no CPU fallback anywhere in this package: if the shared library is
missing, loading fails with an ImportError that says how to build it, and every
compute entry point raises when no B200 is visible.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# NH_B200_LIB points at an alternative build of the same library (instrumented development builds)
LIB_PATH = os.environ.get("NH_B200_LIB") or os.path.join(_HERE, "libnh_b200.so")
CSRC = os.path.join(_HERE, "csrc")

NH_OK, NH_E_SIZE, NH_E_ARG, NH_E_CUDA, NH_E_NOMEM = 0, -1, -2, -3, -4
COST_SAD, COST_SATD = 0, 1

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64

# name -> (restype, argtypes); mirrors include/nh_b200.h declaration by declaration.
PROTOTYPES = {
    "nh_version": (_i, []),
    "nh_last_error": (C.c_char_p, []),
    "nh_device_ok": (_i, []),
    "nh_get_transform_matrix": (_i, [_i, _i, _p]),
    "nh_get_intra_pred_angle": (_i, [_i, C.POINTER(_i)]),
    "nh_get_qp_params": (_i, [_i, C.POINTER(_i), C.POINTER(_i)]),
    "nh_get_quant_scales": (_i, [_i, C.POINTER(_i), C.POINTER(_i)]),
    "nh_forward_transform": (_i, [_p, _i, _p, _i64, _i, _i, _p]),
    "nh_inverse_transform": (_i, [_p, _p, _i64, _i, _i, _p]),
    "nh_quantize": (_i, [_p, _p, _i64, _i, _i, _i, _p]),
    "nh_dequantize": (_i, [_p, _p, _i64, _i, _i, _p]),
    "nh_intra_dc_predict": (_i, [_p, _p, _p, _i64, _i, _p]),
    "nh_intra_dc_predict_ragged": (_i, [_p, _i, _p, _i, _p, _i64, _i, _p]),
    "nh_intra_planar_predict": (_i, [_p, _p, _p, _p, _p, _i64, _i, _p]),
    "nh_intra_predict_modes": (_i, [_p, _p, _p, _p, _i, _i, _p, _i64, _i, _p]),
    "nh_residual_block": (_i, [_p, _p, _p, _i64, _p]),
    "nh_reconstruct_block": (_i, [_p, _p, _p, _i64, _p]),
    "nh_clip_to_pixel_range": (_i, [_p, _p, _i64, _i, _p]),
    "nh_fused_pipeline_dcplanar": (_i, [_p, _p, _p, _p, _p, _p, _i, _i64, _i, _i, _i, _i, _i,
                                        _p, _p, _p, _p, _p]),
    "nh_set_fused_impl": (_i, [_i]),
    "nh_set_rows_impl": (_i, [_i]),
    "nh_set_search_impl": (_i, [_i]),
    "nh_set_wave_impl": (_i, [_i, _i]),
    "nh_fused_pipeline_modes": (_i, [_p, _p, _p, _p, _p, _i, _i64, _i, _i, _i, _i, _i,
                                     _p, _p, _p, _p, _p]),
    "nh_gather_refs": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "nh_plane_to_blocks": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "nh_blocks_to_plane": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "nh_encode_frame_scratch_bytes": (_i64, [_i, _i, _i]),
    "nh_encode_frame": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i64,
                             _p]),
    "nh_encode_frames_scratch_bytes": (_i64, [_i, _i, _i, _i]),
    "nh_encode_frames": (_i, [_p, _i, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p,
                              _i64, _p]),
    "nh_reduce_sse_sad": (_i, [_p, _p, _i64, _p, _p]),
    "nh_reduce_sse_sad_2d": (_i, [_p, _i, _p, _i, _i, _i, _p, _p]),
    "nh_reduce_metrics_i32": (_i, [_p, _p, _i64, _p, _p, _p]),
    "nh_reduce_sse_f64": (_i, [_p, _p, _i64, _p, _p]),
    "nh_satd_4x4_i32": (_i, [_p, _p, _i64, _p, _p]),
    "nh_block_costs": (_i, [_p, _p, _i64, _i, _p, _p, _p, _p]),
    "nh_count_nonzero": (_i, [_p, _i64, _p, _p]),
    "nh_level_stats": (_i, [_p, _i64, _p, _p, _p]),
    "nh_host_encode_frames_scratch_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "nh_host_encode_frames_last_transfer": (_i, [_p, _p]),
    "nh_host_encode_frames": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _i64]),
    "nh_host_pipeline_scratch_bytes": (_i64, [_i, _i64]),
    "nh_host_pipeline_last_transfer": (_i, [_p, _p]),
    "nh_convert_u8_to_i16": (_i, [_p, _p, _i64, _p]),
    "nh_convert_i16_to_u8": (_i, [_p, _p, _i64, _p]),
    "nh_host_pipeline_dcplanar": (_i, [_p, _p, _p, _p, _p, _p, _i, _i64, _i, _i, _i, _i, _i,
                                       _p, _p, _p, _p, _p, _i64, _i64]),
    "nh_host_pipeline_dcplanar_i16": (_i, [_p, _p, _p, _p, _p, _p, _i, _i64, _i, _i, _i, _i, _i,
                                           _p, _p, _p, _p, _p, _i64, _i64]),
}


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a -> nano_hevc_b200/libnh_b200.so (nvcc, in-tree)."""
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libnh_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library with prototypes attached.  Raises ImportError when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA kernel library has not been built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C nano_hevc_b200/csrc` (needs nvcc with sm_100a support). "
                "There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().nh_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a C-ABI return code to the exception the reference would raise."""
    if rc == NH_OK:
        return
    msg = last_error()
    if rc == NH_E_SIZE:
        raise ValueError(msg)  # transform.py:151 raises ValueError("Unsupported transform size: ...")
    if rc == NH_E_ARG:
        raise ValueError(msg)
    if rc == NH_E_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
