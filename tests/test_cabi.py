"""The C-ABI library loads on a CPU-only box and exports every symbol include/advgrpo_b200.h declares.
No compute calls here (no GPU)."""
import ctypes
import os
import re

import pytest

from adv_grpo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "advgrpo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(advgrpo_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    from adv_grpo_b200 import build
    build.build()
    lib = _lib.load()
    assert lib.advgrpo_abi_version() == 1
    assert isinstance(lib.advgrpo_last_error(), bytes)       # empty unless an earlier test in this process hit an error


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/advgrpo_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in adv_grpo_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_bad_arguments_return_error_codes_not_crashes():
    lib = _lib.load()
    # null pointers are rejected before any CUDA call
    with pytest.raises(_lib.AdvGrpoError, match="null"):
        _lib.call("advgrpo_grpo_clip_loss", None, None, None, 1, 4, 1e-5, 5.0, 1.0, None, None, None)
    assert b"null" in lib.advgrpo_last_error()
    with pytest.raises(_lib.AdvGrpoError, match="head_dim"):
        _lib.call("advgrpo_attn_fwd", 16, 16, None, 0, None, 1, 128, 4, 80, 0.1, 0, None)
    assert _lib.query("advgrpo_sde_step_workspace_bytes", 8, 65536) > 0
    with pytest.raises(_lib.AdvGrpoError, match="null"):
        _lib.call("advgrpo_clip_adamw", None, None, None, None, 16, 3e-4, 0.9, 0.999, 1e-8, 1e-4, 1, 1.0, 1, None, None, 0, None)
    with pytest.raises(_lib.AdvGrpoError, match="step"):
        _lib.call("advgrpo_clip_adamw", 16, 16, 16, 16, 16, 3e-4, 0.9, 0.999, 1e-8, 1e-4, 0, 1.0, 1, None, None, 0, None)
    assert _lib.query("advgrpo_clip_adamw_workspace_bytes", 1 << 20) >= 8
    # the host halves of the image decoders: null pointers, empty and one-byte files
    import ctypes
    png_info, jpeg_info = _lib.PngInfo(), _lib.JpegInfo()
    one = (ctypes.c_uint8 * 1)(0xFF)
    for name, args in (("advgrpo_png_parse", (None, 0, ctypes.byref(png_info))), ("advgrpo_png_parse", (one, 1, None)),
                       ("advgrpo_png_parse", (one, 1, ctypes.byref(png_info))), ("advgrpo_jpeg_parse", (None, 0, ctypes.byref(jpeg_info))),
                       ("advgrpo_jpeg_parse", (one, 1, ctypes.byref(jpeg_info))), ("advgrpo_png_inflate", (None, 0, None, None)),
                       ("advgrpo_jpeg_entropy_decode", (None, 0, None, None))):
        with pytest.raises(_lib.AdvGrpoError):
            _lib.call(name, *args)
    for name in ("advgrpo_png_raw_bytes", "advgrpo_png_workspace_bytes", "advgrpo_jpeg_coef_count", "advgrpo_jpeg_workspace_bytes"):
        assert _lib.query(name, None) == 0


def test_ops_refuse_cpu_tensors():
    import torch
    from adv_grpo_b200 import ops
    with pytest.raises(_lib.AdvGrpoError, match="CUDA"):
        ops.group_advantage(torch.zeros(4), torch.zeros(4, dtype=torch.int64))
    from adv_grpo_b200.optim import FlatClipAdamW, TorchOrderAdam
    with pytest.raises(ValueError, match="CUDA"):
        FlatClipAdamW([torch.nn.Parameter(torch.zeros(8))])
    with pytest.raises(ValueError, match="CUDA"):
        TorchOrderAdam([torch.nn.Parameter(torch.zeros(8))])
    # the discriminator-step / score-head ops have no CPU path either
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    for call in (lambda: ops.col_sum(x), lambda: ops.gemm_tn(x, x), lambda: ops.linear(x, x),
                 lambda: ops.attention_small(x.view(1, 4, 1, 64), x.view(1, 4, 1, 64), x.view(1, 4, 1, 64)),
                 lambda: ops.gather_rows_l2norm(x.view(1, 4, 64), torch.zeros(1, 2, dtype=torch.int64), True),
                 lambda: ops.pil_resize_bilinear(torch.zeros(8, 8, 3, dtype=torch.uint8), 4, 4),
                 lambda: ops.row_softmax_f32(torch.zeros(2, 8))):
        with pytest.raises(_lib.AdvGrpoError, match="CUDA"):
            call()


def test_empty_inputs_are_noops_or_clean_errors():
    """Edge of every size range, checked without a device (the entry points decide before any CUDA call): the
    streaming ops accept an empty batch as a no-op returning 0; the tensor-core / image ops refuse empty shapes with
    ADVGRPO_ERR_BAD_ARG and a message instead of launching a zero-sized grid."""
    buf = 4096                     # a non-null, 16-byte aligned address that is never dereferenced for empty inputs
    ok = [
        ("advgrpo_group_advantage", (buf, buf, 1, 0, 2, 1, buf, None, None, 0, None)),
        ("advgrpo_group_advantage_mode", (buf, buf, 1, 0, 1, 0, 3, buf, None, None, 0, None)),
        ("advgrpo_ln_modulate_fwd", (buf, buf, buf, None, None, 1536, buf, None, 0, 77, 1536, 1e-6, None)),
        ("advgrpo_ln_modulate_bwd", (buf, buf, None, 1536, buf, None, buf, 0, 0, 77, 1536, 1e-6, None)),
        ("advgrpo_layer_norm_affine", (buf, buf, buf, buf, 0, 1280, 1e-5, None)),
        ("advgrpo_row_gate_mul", (buf, buf, 1536, 1, buf, 0, 1536, None)),
        ("advgrpo_qk_norm_concat_fwd", (buf, None, buf, buf, None, None, buf, 0, 16, 0, 24, 64, 1e-6, None)),
        ("advgrpo_clip_adamw", (buf, buf, buf, buf, 0, 3e-4, 0.9, 0.999, 1e-8, 1e-4, 1, 1.0, 1, None, None, 0, None)),
        # score heads / discriminator step (csrc/heads.cu): an empty batch of images / rows / parameters is a no-op
        ("advgrpo_gather_rows_l2norm", (buf, buf, buf, 0, 1370, 64, 768, 1, 1e-6, None)),
        ("advgrpo_head_logits", (buf, buf, buf, buf, 0, 512, 1, None)),
        ("advgrpo_dino_hybrid_score", (buf, buf, 0, 64, 0.7, 1, None)),
        ("advgrpo_head_dz", (buf, buf, buf, buf, 0, 512, None)),
        ("advgrpo_pickscore_head", (buf, buf, None, buf, 1, buf, 0, 1, 1024, 0, None)),
        ("advgrpo_adam_torch_order", (buf, buf, buf, buf, 0, 1, 1, 5e-6, 0.5, 0.999, 1e-8, 1, 0, None)),
        ("advgrpo_row_softmax_f32", (buf, buf, 0, 4096, 1.0, 1, None)),
        ("advgrpo_col_sum", (buf, 512, None, 0, None, buf, 0, 0, None, 0, None)),
    ]
    for name, args in ok:
        assert _lib.call(name, *args) == 0, name
    bad = [
        ("advgrpo_attn_fwd", (buf, buf, None, 0, None, 0, 128, 4, 64, 0.125, 0, None)),
        ("advgrpo_attn_bwd", (buf, buf, buf, buf, buf, 0, 128, 4, 64, 0.125, 0, buf, 1 << 20, None)),
        ("advgrpo_conv2d_nhwc_tf32", (buf, buf, None, buf, 0, 8, 8, 64, 64, 3, None)),
        ("advgrpo_dino_preprocess", (buf, 0, 0, 64, 64, 518, buf, buf, buf, None)),
        ("advgrpo_upsample_nearest2x_nhwc", (buf, buf, 0, 8, 8, 64, None)),
        ("advgrpo_gather_rows_l2norm", (buf, buf, buf, 2, 1370, 64, 770, 1, 1e-6, None)),            # D not a multiple of 8
        ("advgrpo_dino_hinge_loss", (buf, buf, buf, 0, 4, 64, 0.3, None)),                           # no real images
        ("advgrpo_adam_torch_order", (buf, buf, buf, buf, 16, 1, 1, 5e-6, 0.5, 0.999, 1e-8, 0, 0, None)),   # step 0
        ("advgrpo_layer_norm_affine_bwd", (buf, buf, buf, buf, buf, None, 4, 1280, 1e-5, None, 0, None)),   # dweight without dbias
        ("advgrpo_attn_small_fwd", (buf, buf, buf, buf, None, 1, 257, 16, 81, 0.1, 0, None)),        # odd head_dim
        ("advgrpo_attn_small_bwd", (buf, buf, buf, buf, buf, buf, buf, buf, buf, buf, 1, 4096, 1, 64, 0.125, 0, None)),  # too long
        ("advgrpo_row_softmax_f32", (buf, buf, 3, 6, 1.0, 0, None)),                                 # cols not a multiple of 4
        ("advgrpo_jpeg_parse", (None, 0, None)),                                                     # null pointers
        ("advgrpo_png_parse", (None, 0, None)),
        ("advgrpo_png_unfilter_to_rgb", (buf, buf, None, buf, buf, 1 << 20, None)),                  # no header info
        ("advgrpo_jpeg_idct_to_rgb", (buf, buf, None, buf, buf, 1 << 20, None)),                     # no frame info
        ("advgrpo_pil_resize_bilinear_u8", (buf, 0, 640, 512, 512, buf, None, buf, 1 << 20, None)),  # empty image
        ("advgrpo_pil_resize_bilinear_u8", (buf, 480, 640, 512, 512, buf, None, buf, 16, None)),     # workspace too small
    ]
    for name, args in bad:
        with pytest.raises(_lib.AdvGrpoError):
            _lib.call(name, *args)
        assert len(_lib.load().advgrpo_last_error()) > 0
