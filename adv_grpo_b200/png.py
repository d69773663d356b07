"""PNG files -> RGB bytes on the device: the adversarial loop's reference images are PNG files
(README.md:114-128 of the reference) that the scripts open with `Image.open(fpath).convert("RGB")` on the host
(`scripts/train_sd3_fast_pickscore.py:773-786`).  Here the IDAT stream is inflated by the library's own host inflate
(`advgrpo_png_inflate`, plain C++), the filtered scan lines go to the GPU, and unfiltering (a wavefront over the image's
anti-diagonals; once per reduced image of an Adam7-interlaced file) and the conversion to RGB run there (`csrc/png.cu`),
byte-exact with Pillow.  Files the decoder does not take (malformed headers) return None: the caller hands them to Pillow, which reports the error."""
import ctypes

import torch

from . import _lib, ops


def png_info(data):
    buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)
    info = _lib.PngInfo()
    _lib.call("advgrpo_png_parse", buf, len(data), ctypes.byref(info))
    return info


def inflate(data, info=None):
    """Host step: -> (filtered scan lines uint8 [H * (1 + rowbytes)], palette uint8 [768], info); pinned tensors with CUDA."""
    info = info or png_info(data)
    if not info.supported:
        return None, None, info
    pin = torch.cuda.is_available()
    raw = torch.empty(_lib.query("advgrpo_png_raw_bytes", ctypes.byref(info)), dtype=torch.uint8, pin_memory=pin)
    pal = torch.empty(768, dtype=torch.uint8, pin_memory=pin)
    buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)
    _lib.call("advgrpo_png_inflate", buf, len(data), raw.data_ptr(), pal.data_ptr())
    return raw, pal, info


def decode_png_to_device(data, device="cuda"):
    """bytes of a PNG file -> uint8 [H, W, 3] on `device`, equal to `np.asarray(Image.open(...).convert("RGB"))`; None when the
    file is outside the supported subset."""
    raw, pal, info = inflate(data)
    if raw is None:
        return None
    return unfilter_on_device(raw, pal, info, device)


def unfilter_on_device(raw, pal, info, device="cuda"):
    """Device step: the output of `inflate` (host tensors) -> uint8 [H, W, 3] on `device`."""
    dev = torch.device(device)
    raw_d, pal_d = raw.to(dev, non_blocking=True), pal.to(dev, non_blocking=True)
    rgb = torch.empty((info.height, info.width, 3), dtype=torch.uint8, device=dev)
    ws = ops._workspace("png", _lib.query("advgrpo_png_workspace_bytes", ctypes.byref(info)), dev)
    with torch.cuda.device(dev):
        _lib.call("advgrpo_png_unfilter_to_rgb", raw_d.data_ptr(), pal_d.data_ptr(), ctypes.byref(info), rgb.data_ptr(),
                  ws.data_ptr(), ws.numel(), ops._stream())
    return rgb
