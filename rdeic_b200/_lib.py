"""ctypes binding of librdeic_b200.so (declared in include/rdeic_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
product path raises.  `python -m rdeic_b200.build` (or `__graft_entry__.build()`) produces the
library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
import os as _os

# RDEIC_B200_LIB selects an alternative build of the same ABI (kernel A/B experiments)
LIB_PATH = Path(_os.environ.get("RDEIC_B200_LIB", _HERE / "librdeic_b200.so"))

c_f32p = C.c_void_p
vp = C.c_void_p
i32 = C.c_int
i64 = C.c_int64
f32 = C.c_float


class ConvParams(C.Structure):
    """Mirror of `rdeic_conv_params` (include/rdeic_b200.h)."""

    _fields_ = [
        ("a", vp), ("a_n", i32), ("a_h", i32), ("a_w", i32), ("a_c", i32),
        ("a2", vp), ("a2_c", i32),
        ("taps", i32),
        ("w", vp),
        ("w_batch_stride", i64),
        ("w_k", i32), ("w_ld", i32),
        ("n_out", i32),
        ("bias", vp),
        ("row_bias", vp), ("row_bias_ld", i32),
        ("resid", vp), ("resid_is_f32", i32), ("ld_resid", i32),
        ("alpha", f32),
        ("act", i32),
        ("out_bf16", vp),
        ("out_f32", vp),
        ("ldo", i32),
        ("tile_n_hint", i32),
        ("workspace", vp),
        ("workspace_bytes", i64),
        ("act_param", f32),
        ("a_ld", i32), ("a2_ld", i32),
        ("stats_out", vp),
        ("up2", i32), ("a2_center", i32), ("w2", vp), ("in_stride2", i32), ("pad_lo", i32),
    ]


class ConvF32Params(C.Structure):
    """Mirror of `rdeic_conv_f32_params` (fp32 kernel mode)."""

    _fields_ = [
        ("a", vp), ("a_n", i32), ("a_h", i32), ("a_w", i32), ("c1", i32),
        ("a2", vp), ("c2", i32),
        ("ksize", i32), ("stride", i32), ("up", i32),
        ("w", vp),
        ("n_out", i32),
        ("bias", vp),
        ("row_bias", vp), ("row_bias_ld", i32),
        ("resid", vp), ("ld_resid", i32),
        ("alpha", f32),
        ("act", i32),
        ("out", vp), ("ldo", i32),
        ("act_param", f32),
        ("a_ld", i32), ("a2_ld", i32),
        ("workspace", vp), ("workspace_bytes", i64),
    ]


# name -> argtypes (return type is int unless listed in _RESTYPES)
SIGNATURES = {
    "rdeic_last_error": [],
    "rdeic_abi_version": [],
    "rdeic_ckbd_mask": [vp, vp, i32, i32, i32, i32, i32, vp],
    "rdeic_ckbd_split": [vp, vp, vp, i32, i32, i32, i32, vp],
    "rdeic_ckbd_merge": [vp, vp, vp, i64, vp],
    "rdeic_ckbd_squeeze": [vp, vp, i32, i32, i32, i32, i32, vp],
    "rdeic_ckbd_unsqueeze": [vp, vp, i32, i32, i32, i32, i32, vp],
    "rdeic_quantize_symbols": [vp, vp, vp, i64, vp],
    "rdeic_dequantize": [vp, vp, vp, i64, vp],
    "rdeic_build_indexes": [vp, vp, i32, f32, vp, i64, vp],
    "rdeic_ckbd_squeeze_indexes": [vp, vp, vp, i32, f32, vp, vp, i32, i32, i32, i32, i32, vp],
    "rdeic_ckbd_encode_phase": [vp, vp, vp, vp, i32, f32, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "rdeic_ckbd_decode_phase": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "rdeic_vq_quant": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "rdeic_vq_lookup": [vp, vp, vp, i32, i32, i32, i32, vp],
    "rdeic_q_sample": [vp, vp, vp, i64, f32, f32, vp],
    "rdeic_relay_update": [vp, vp, vp, f32, vp, vp, i64, f32, f32, f32, f32, f32, vp],
    "rdeic_ddim_update": [vp, vp, vp, f32, vp, vp, vp, i64, f32, f32, f32, f32, f32, vp],
    "rdeic_nchw_to_nhwc_bf16": [vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "rdeic_nhwc_to_nchw_f32": [vp, i32, vp, i32, i32, i32, i32, i32, vp],
    "rdeic_f32_to_bf16": [vp, vp, i64, vp],
    "rdeic_timestep_embedding": [vp, vp, i32, i32, f32, vp],
    "rdeic_silu_bf16": [vp, i32, vp, i64, vp],
    "rdeic_geglu": [vp, vp, i64, i32, vp],
    "rdeic_upsample2x_nhwc": [vp, vp, i32, i32, i32, i32, vp],
    "rdeic_pixel_shuffle2_nhwc": [vp, vp, i32, i32, i32, i32, vp],
    "rdeic_im2col_3x3_s2": [vp, vp, i32, i32, i32, i32, i32, vp],
    "rdeic_softmax_rows": [vp, i32, vp, i64, i32, f32, vp],
    "rdeic_transpose_bf16": [vp, vp, i32, i32, i32, vp],
    "rdeic_image_to_u8": [vp, vp, i64, i32, vp],
    "rdeic_split_bf16_hilo": [vp, i64, i32, i64, vp, i64, i32, i32, vp],
    "rdeic_blend_tiles_u8": [vp, vp, i32, i32, i32, i32, vp, i32, i32, vp],
    "rdeic_groupnorm_workspace_bytes": [i32, i64, i32],
    "rdeic_groupnorm_is_small": [i32, i64, i32, i32, i32],
    "rdeic_groupnorm_nhwc": [vp, i32, vp, i32, i32, vp, vp, vp, i32, i64, i32, f32, i32, vp, vp],
    "rdeic_groupnorm_from_stats": [vp, i32, vp, vp, i32, vp, i32, vp, vp, vp, i32, i64, i32, f32, i32, vp, vp],
    "rdeic_gn_silu_conv3x3_tail": [vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, f32, vp, vp],
    "rdeic_layernorm": [vp, i32, vp, vp, vp, i64, i32, f32, vp],
    "rdeic_conv_gemm": [C.POINTER(ConvParams), vp],
    "rdeic_conv_stats_supported": [i32, i32, i32, i32, i32],
    "rdeic_pack_conv_weight": [vp, vp, i32, i32, i32, i32, i32, vp],
    "rdeic_conv_f32": [C.POINTER(ConvF32Params), vp],
    "rdeic_attention_f32": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, f32, vp],
    "rdeic_geglu_f32": [vp, vp, i64, i32, vp],
    "rdeic_timestep_embedding_f32": [vp, vp, i32, i32, f32, vp],
    "rdeic_groupnorm_nhwc_f32": [vp, i32, vp, i32, vp, vp, vp, i32, i64, i32, f32, i32, vp, vp],
    "rdeic_layernorm_f32": [vp, vp, vp, vp, i64, i32, f32, vp],
    "rdeic_attention": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, f32, vp],
}
_RESTYPES = {"rdeic_last_error": C.c_char_p, "rdeic_groupnorm_workspace_bytes": i64}

_lib = None


class RdeicLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once) and attach prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RdeicLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m rdeic_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path."
        )
    lib = C.CDLL(str(LIB_PATH))
    partial = _os.environ.get("RDEIC_B200_LIB_PARTIAL") is not None    # A/B runs against an older build (scripts/ab_gemm.py)
    for name, argtypes in SIGNATURES.items():
        if partial and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().rdeic_last_error().decode("utf-8", "replace")
        raise RdeicLibraryError(f"{what} failed: {msg}")
