"""JPEG files (sequential and progressive Huffman) -> RGB bytes on the device: the reference opens every reference image with
`Image.open(fpath).convert("RGB")` on the host (`scripts/train_sd3_fast_pickscore.py:773-786`).  Here the file bytes are
entropy-decoded by the library's own host Huffman decoder (`advgrpo_jpeg_entropy_decode`, plain C++), the coefficient
blocks go to the GPU, and the inverse DCT, chroma upsampling and colour conversion run there (`csrc/jpeg.cu`), bit-exact
with libjpeg / Pillow.  Files outside the supported subset (arithmetic-coded, CMYK, ...) return None: the caller keeps Pillow for
them -- an explicit, per-file host decode, not a silent fallback of a kernel."""
import ctypes

import numpy as np
import torch

from . import _lib, ops


def jpeg_info(data):
    """Parsed frame header (`_lib.JpegInfo`); `.supported == 0` for a valid JPEG this decoder does not take."""
    buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)
    info = _lib.JpegInfo()
    _lib.call("advgrpo_jpeg_parse", buf, len(data), ctypes.byref(info))
    return info


def entropy_decode(data, info=None):
    """Host step: -> (coefs int16 [n] as a pinned tensor when CUDA is available, qtabs uint16-as-int16 [192], info)."""
    info = info or jpeg_info(data)
    if not info.supported:
        return None, None, info
    n = _lib.query("advgrpo_jpeg_coef_count", ctypes.byref(info))
    pin = torch.cuda.is_available()
    coefs = torch.empty(n, dtype=torch.int16, pin_memory=pin)
    qt = torch.empty(192, dtype=torch.int16, pin_memory=pin)
    buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)
    _lib.call("advgrpo_jpeg_entropy_decode", buf, len(data), coefs.data_ptr(), qt.data_ptr())
    return coefs, qt, info


def decode_jpeg_to_device(data, device="cuda"):
    """bytes of a JPEG file -> uint8 [H, W, 3] on `device`, equal to `np.asarray(Image.open(...).convert("RGB"))`; None when the
    file is outside the supported subset."""
    coefs, qt, info = entropy_decode(data)
    if coefs is None:
        return None
    return idct_on_device(coefs, qt, info, device)


def idct_on_device(coefs, qt, info, device="cuda"):
    """Device step: the output of `entropy_decode` (host tensors) -> uint8 [H, W, 3] on `device`."""
    dev = torch.device(device)
    coefs_d, qt_d = coefs.to(dev, non_blocking=True), qt.to(dev, non_blocking=True)
    rgb = torch.empty((info.height, info.width, 3), dtype=torch.uint8, device=dev)
    ws = ops._workspace("jpeg", _lib.query("advgrpo_jpeg_workspace_bytes", ctypes.byref(info)), dev)
    with torch.cuda.device(dev):
        _lib.call("advgrpo_jpeg_idct_to_rgb", coefs_d.data_ptr(), qt_d.data_ptr(), ctypes.byref(info), rgb.data_ptr(),
                  ws.data_ptr(), ws.numel(), ops._stream())
    return rgb


def coefficients_as_numpy(data):
    """Test hook: per-component int16 [blocks_h, blocks_w, 64] arrays of the host entropy decoder."""
    coefs, qt, info = entropy_decode(data)
    if coefs is None:
        return None, None, info
    out, off = [], 0
    arr = coefs.numpy()
    for c in range(info.ncomp):
        k = info.blocks_w[c] * info.blocks_h[c] * 64
        out.append(arr[off:off + k].reshape(info.blocks_h[c], info.blocks_w[c], 64).copy())
        off += k
    return out, qt.numpy().view(np.uint16).reshape(3, 64).copy(), info
