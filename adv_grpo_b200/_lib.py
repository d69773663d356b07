"""ctypes binding of libadvgrpo_b200.so (the C ABI declared in include/advgrpo_b200.h).

There is no fallback: if the shared object is missing or a call fails, an exception is
raised -- the product path never silently degrades to PyTorch/CPU arithmetic.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libadvgrpo_b200.so")

_P, _I64, _I, _F, _D, _U64, _SZ = c_void_p, c_int64, c_int, c_float, c_double, c_uint64, c_size_t

# name -> (restype, argtypes); mirrors include/advgrpo_b200.h one to one.
SIGNATURES = {
    "advgrpo_abi_version": (c_int, []),
    "advgrpo_last_error": (c_char_p, []),
    "advgrpo_device_check": (c_int, [_I]),
    "advgrpo_sde_step_workspace_bytes": (_SZ, [_I64, _I64]),
    "advgrpo_cfg_sde_step_logprob": (c_int, [_P, _P, _P, _P, _P, _P, _I64, _P, _P, _I64, _P, _P, _P, _P,
                                             _I64, _I64, _F, _F, _U64, _U64, _P, _SZ, _P]),
    "advgrpo_cfg_sde_logprob_bwd": (c_int, [_P, _P, _P, _P, _P, _I64, _P, _P, _I64, _P, _P, _P, _I64,
                                            _I64, _F, _F, _P]),
    "advgrpo_cfg_sde_logprob_kl_bwd": (c_int, [_P, _P, _P, _P, _P, _I64, _P, _P, _I64, _P, _P, _P, _P, _P, _I64,
                                               _I64, _F, _F, _P]),
    "advgrpo_cfg_sde_step_logprob_variant": (c_int, [_P, _P, _P, _P, _P, _P, _I64, _P, _P, _I64, _P, _P, _P, _P,
                                                     _I64, _I64, _F, _F, _U64, _U64, _P, _SZ, _I, _P]),
    "advgrpo_cfg_sde_logprob_bwd_variant": (c_int, [_P, _P, _P, _P, _P, _I64, _P, _P, _I64, _P, _P, _P, _P, _P, _I64,
                                                    _I64, _F, _F, _I, _P]),
    "advgrpo_group_advantage_workspace_bytes": (_SZ, [_I64, _I64]),
    "advgrpo_group_advantage": (c_int, [_P, _P, _I64, _I64, _I64, _I, _P, _P, _P, _SZ, _P]),
    "advgrpo_group_advantage_mode": (c_int, [_P, _P, _I64, _I64, _I64, _I, _I, _P, _P, _P, _SZ, _P]),
    "advgrpo_grpo_clip_loss": (c_int, [_P, _P, _P, _I64, _I64, _D, _D, _D, _P, _P, _P]),
    "advgrpo_clip_adamw_workspace_bytes": (_SZ, [_I64]),
    "advgrpo_clip_adamw": (c_int, [_P, _P, _P, _P, _I64, _D, _D, _D, _D, _D, _I64, _D, _I, _P, _P, _SZ, _P]),
    "advgrpo_ln_modulate_fwd": (c_int, [_P, _P, _P, _P, _P, _I64, _P, _P, _I64, _I64, _I64, _F, _P]),
    "advgrpo_ln_modulate_bwd": (c_int, [_P, _P, _P, _I64, _P, _P, _P, _I, _I64, _I64, _I64, _F, _P]),
    "advgrpo_layer_norm_affine": (c_int, [_P, _P, _P, _P, _I64, _I64, _F, _P]),
    "advgrpo_qk_norm_concat_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _F, _P]),
    "advgrpo_qk_norm_concat_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64,
                                           _I64, _F, _P]),
    "advgrpo_attn_fwd": (c_int, [_P, _P, _P, _I64, _P, _I64, _I64, _I64, _I64, _F, _I, _P]),
    "advgrpo_attn_fwd_bias": (c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _F, _P]),
    "advgrpo_attn_bwd_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I64]),
    "advgrpo_attn_bwd": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _F, _I, _P, _SZ, _P]),
    "advgrpo_gemm_bf16": (c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _I64, _P, _P, _I64, _I64, _I64,
                                  _I64, _I, _P, _I64, _P, _I64, _I64, _P, _P]),
    "advgrpo_gemm_bf16_dual": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _I64, _I64, _I, _P, _P, _P,
                                       _P, _P, _P, _P]),
    "advgrpo_row_gate_mul": (c_int, [_P, _P, _I64, _I64, _P, _I64, _I64, _P]),
    "advgrpo_gemm_tn_skinny_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "advgrpo_gemm_tn_skinny": (c_int, [_P, _P, _P, _I64, _I64, _I64, _I, _P, _SZ, _P]),
    "advgrpo_gemm_qkv_norm": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _I64, _P, _I64, _I64, _I64,
                                      _F, _P]),
    "advgrpo_clip_preprocess_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I64]),
    "advgrpo_clip_preprocess": (c_int, [_P, _I, _I64, _I64, _I64, _I64, _P, _P, _P, _I, _P, _P, _SZ, _P]),
    "advgrpo_dino_preprocess": (c_int, [_P, _I, _I64, _I64, _I64, _I64, _P, _P, _P, _P]),
    "advgrpo_jpeg_parse": (c_int, [_P, _SZ, _P]),
    "advgrpo_jpeg_coef_count": (_SZ, [_P]),
    "advgrpo_jpeg_entropy_decode": (c_int, [_P, _SZ, _P, _P]),
    "advgrpo_jpeg_workspace_bytes": (_SZ, [_P]),
    "advgrpo_jpeg_idct_to_rgb": (c_int, [_P, _P, _P, _P, _P, _SZ, _P]),
    "advgrpo_png_parse": (c_int, [_P, _SZ, _P]),
    "advgrpo_png_raw_bytes": (_SZ, [_P]),
    "advgrpo_png_inflate": (c_int, [_P, _SZ, _P, _P]),
    "advgrpo_png_workspace_bytes": (_SZ, [_P]),
    "advgrpo_png_unfilter_to_rgb": (c_int, [_P, _P, _P, _P, _P, _SZ, _P]),
    "advgrpo_pil_resize_bilinear_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I64]),
    "advgrpo_pil_resize_bilinear_u8": (c_int, [_P, _I64, _I64, _I64, _I64, _P, _P, _P, _SZ, _P]),
    "advgrpo_group_norm_workspace_bytes": (_SZ, [_I64, _I64]),
    "advgrpo_group_norm_silu_nhwc": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _F, _I, _P, _SZ, _P]),
    "advgrpo_conv2d_nhwc_tf32": (c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I, _P]),
    "advgrpo_add_bias_nhwc": (c_int, [_P, _P, _P, _P, _I64, _I64, _P]),
    "advgrpo_upsample_nearest2x_nhwc": (c_int, [_P, _P, _I64, _I64, _I64, _I64, _P]),
    "advgrpo_gather_rows_l2norm": (c_int, [_P, _P, _P, _I64, _I64, _I64, _I64, _I, _F, _P]),
    "advgrpo_head_logits": (c_int, [_P, _P, _P, _P, _I64, _I64, _I, _P]),
    "advgrpo_dino_hybrid_score": (c_int, [_P, _P, _I64, _I64, _F, _I, _P]),
    "advgrpo_dino_hinge_loss": (c_int, [_P, _P, _P, _I64, _I64, _I64, _F, _P]),
    "advgrpo_head_dz": (c_int, [_P, _P, _P, _P, _I64, _I64, _P]),
    "advgrpo_col_sum_workspace_bytes": (_SZ, [_I64, _I64]),
    "advgrpo_col_sum": (c_int, [_P, _I64, _P, _I64, _P, _P, _I64, _I64, _P, _SZ, _P]),
    "advgrpo_pickscore_head": (c_int, [_P, _P, _P, _P, _I, _P, _I64, _I64, _I64, _I, _P]),
    "advgrpo_layer_norm_affine_bwd_workspace_bytes": (_SZ, [_I64, _I64]),
    "advgrpo_layer_norm_affine_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _I64, _I64, _F, _P, _SZ, _P]),
    "advgrpo_ln_modulation_grads_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "advgrpo_ln_modulation_grads": (c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _F, _P, _SZ, _P]),
    "advgrpo_adam_torch_order": (c_int, [_P, _P, _P, _P, _I64, _I, _I, _D, _D, _D, _D, _I64, _I, _P]),
    "advgrpo_row_softmax_f32": (c_int, [_P, _P, _I64, _I64, _F, _I, _P]),
    "advgrpo_attn_small_fwd": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _F, _I, _P]),
    "advgrpo_attn_small_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _F, _I, _P]),
}
# test/bench hooks that are exported but not part of include/advgrpo_b200.h
_EXTRA = {
    "advgrpo_attn_fwd_variant": (c_int, [_P, _P, _P, _I64, _I64, _I64, _I64, _F, _I, _I, _P]),
    "advgrpo_debug_set_gemm_variant": (None, [_I]),
    "advgrpo_debug_set_conv_variant": (None, [_I]),
    "advgrpo_debug_set_pdl": (None, [_I]),
    "advgrpo_debug_set_attn_trace": (None, [_P]),
}

class JpegInfo(ctypes.Structure):
    """`advgrpo_jpeg_info` of include/advgrpo_b200.h."""
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("ncomp", ctypes.c_int32),
                ("h", ctypes.c_int32 * 3), ("v", ctypes.c_int32 * 3), ("tq", ctypes.c_int32 * 3),
                ("blocks_w", ctypes.c_int32 * 3), ("blocks_h", ctypes.c_int32 * 3),
                ("restart_interval", ctypes.c_int32), ("supported", ctypes.c_int32), ("progressive", ctypes.c_int32)]


class PngInfo(ctypes.Structure):
    """`advgrpo_png_info` of include/advgrpo_b200.h."""
    _fields_ = [(n, ctypes.c_int32) for n in ("width", "height", "bit_depth", "color_type", "interlace", "channels", "rowbytes",
                                              "palette_entries", "supported")]


_lib = None


class AdvGrpoError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m adv_grpo_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in {**SIGNATURES, **_EXTRA}.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.advgrpo_abi_version() != 1:
        raise ImportError("libadvgrpo_b200.so ABI version mismatch")
    _lib = lib
    return lib


# kernels launched per successful entry-point call (bench.py's `gpu_launches` claim)
_KERNELS_PER_CALL = {"advgrpo_group_norm_silu_nhwc": 2, "advgrpo_attn_bwd": 3, "advgrpo_clip_preprocess": 3, "advgrpo_group_advantage": 2, "advgrpo_group_advantage_mode": 2, "advgrpo_clip_adamw": 2,
                     "advgrpo_col_sum": 2, "advgrpo_layer_norm_affine_bwd": 3, "advgrpo_attn_small_bwd": 2, "advgrpo_ln_modulation_grads": 3, "advgrpo_pil_resize_bilinear_u8": 4, "advgrpo_jpeg_parse": 0, "advgrpo_jpeg_entropy_decode": 0,
                     "advgrpo_jpeg_idct_to_rgb": 4, "advgrpo_png_parse": 0, "advgrpo_png_inflate": 0, "advgrpo_png_unfilter_to_rgb": 2, "advgrpo_device_check": 0}
_launches = [0]


def launch_count():
    return _launches[0]


def add_launches(n):
    """CUDA-graph replays re-launch the kernels recorded at capture time."""
    _launches[0] += n


def call(name, *args):
    """Invoke an int-returning entry point; raise with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    _launches[0] += _KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        msg = lib.advgrpo_last_error().decode("utf-8", "replace")
        raise AdvGrpoError(f"{name} failed ({rc}): {msg}")
    return rc


def query(name, *args):
    """Invoke a size_t-returning helper."""
    return getattr(load(), name)(*args)
