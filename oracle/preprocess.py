"""Oracle: reward-image preprocessing.

* `quantise_bf16` -- `(images * 255).round().clamp(0, 255).to(torch.uint8)` on the bf16 image
  tensor (`adv_grpo/rewards.py:581`, images arrive as bf16 from
  `scripts/train_sd3_fast_pickscore.py:816`; quirk Q6).
* `pil_bicubic_resize_u8` -- numpy restatement of Pillow's antialiased BICUBIC resize for 8-bit
  images (`ImagingResample`: precompute_coeffs / normalize_coeffs_8bpc / horizontal then vertical
  pass), which is what `CLIPProcessor` runs (`adv_grpo/pickscore_scorer.py:21-28`).  Pillow is
  third-party; `tests/test_oracle_models.py` checks this restatement bit-exactly against the
  installed Pillow.
* `clip_pixel_values` -- rescale 1/255 + normalise (transformers CLIPImageProcessor).
Test infrastructure only (see oracle/__init__.py)."""
import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)


def quantise_bf16(images_bf16):
    """`(images * 255).round().clamp(0, 255).to(torch.uint8)` on the bf16 images (rewards.py:581)."""
    return (images_bf16 * 255).round().clamp(0, 255).to(torch.uint8)


def _bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Pillow ImagingResample precompute_coeffs (bicubic, antialiased), the resize CLIPProcessor applies at pickscore_scorer.py:21-28."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, kk = [], np.zeros((out_size, ksize), dtype=np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w[1:], w[0]) if w else 0.0
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            s = v * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + s) if v < 0 else int(0.5 + s)
        bounds.append((xmin, xmax))
    return bounds, kk


def _resample_axis_last(img_u8, out_size):
    """img_u8: [..., in] uint8 -> [..., out] uint8 along the last axis."""
    in_size = img_u8.shape[-1]
    bounds, kk = precompute_coeffs(in_size, out_size)
    out = np.empty(img_u8.shape[:-1] + (out_size,), dtype=np.uint8)
    src = img_u8.astype(np.int64)
    for xx, (xmin, xmax) in enumerate(bounds):
        ss = (1 << (PRECISION_BITS - 1)) + (src[..., xmin:xmin + xmax] * kk[xx, :xmax]).sum(-1)
        out[..., xx] = np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_bicubic_resize_u8(img_u8, out_size):
    """img_u8: [..., H, W] uint8 planes -> [..., out, out]; horizontal pass, then vertical."""
    h = _resample_axis_last(img_u8, out_size)
    v = _resample_axis_last(np.swapaxes(h, -1, -2), out_size)
    return np.swapaxes(v, -1, -2)


def clip_pixel_values(u8_chw):
    """CLIPProcessor rescale (1/255) + normalise after the resize / centre crop (pickscore_scorer.py:21-28).
    [B,3,h,w] uint8 -> float32 normalised (rescale in float64, cast f32, normalise in f32)."""
    f = (u8_chw.astype(np.float64) * (1 / 255)).astype(np.float32)
    return (f - CLIP_MEAN[None, :, None, None]) / CLIP_STD[None, :, None, None]
