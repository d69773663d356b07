"""Synthetic PNG files for the decoder tests: Pillow-written ones (its encoder picks None / Sub / Up / Paeth per row) and
hand-assembled ones (any filter type per row incl. Average, any deflate level incl. stored blocks, IDAT split into pieces)."""
import io
import struct
import zlib

import numpy as np

CH = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}


def _chunk(t, b):
    return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xFFFFFFFF)


ADAM7 = ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2))


def handmade_png(h, w, ct, seed=0, level=6, split=None, kind=0, bd=8, interlace=0):
    """A PNG whose rows carry random filter types (the pixels are whatever the filters reconstruct); with interlace=1 the seven
    Adam7 reduced images each get their own scan lines.  Returns (file bytes, the filtered scan-line stream)."""
    rng = np.random.default_rng(seed)
    passes = [(w, h)] if not interlace else [(-(-(w - x0) // dx), -(-(h - y0) // dy)) for x0, y0, dx, dy in ADAM7]
    raw = b""
    for pw, ph in passes:
        if pw <= 0 or ph <= 0:
            continue
        rb = (pw * CH[ct] * bd + 7) // 8
        if kind == 0:
            body = rng.integers(0, 256, (ph, rb), dtype=np.uint8)
        elif kind == 1:
            body = (np.add.outer(np.arange(ph) * 3, np.arange(rb) * 2) % 256).astype(np.uint8)
        else:
            body = np.full((ph, rb), 7, np.uint8)
        ft = rng.integers(0, 5, (ph, 1), dtype=np.uint8)
        raw += np.concatenate([ft, body], 1).tobytes()
    comp = zlib.compress(raw, level)
    idats = [comp] if not split else [comp[i:i + split] for i in range(0, len(comp), split)]
    plte = _chunk(b"PLTE", rng.integers(0, 256, 768, dtype=np.uint8).tobytes()) if ct == 3 else b""
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bd, ct, 0, 0, interlace)) + plte +
            b"".join(_chunk(b"IDAT", c) for c in idats) + _chunk(b"IEND", b"")), raw


def pillow_png(h, w, mode="RGB", seed=0, **kw):
    from PIL import Image
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 9.0 + yy / 13.0), 128 + 90 * np.cos(xx / 7.0), 128 + 80 * np.sin(yy / 5.0),
                    200 + 50 * np.sin(xx / 3.0)], -1) + rng.normal(0, 6, (h, w, 4))
    img = np.clip(img, 0, 255).astype(np.uint8)
    n = {"RGB": 3, "RGBA": 4, "L": 1, "LA": 2, "P": 3}[mode]
    im = Image.fromarray(img[..., :n] if n > 1 else img[..., 0], mode="RGB" if mode == "P" else mode)
    if mode == "P":
        im = im.quantize(200)
    buf = io.BytesIO()
    im.save(buf, format="PNG", **kw)
    return buf.getvalue()
