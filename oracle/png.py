"""Oracle (test infrastructure only): PNG decode restated in numpy -- what Pillow's `Image.open(path).convert("RGB")`
(`scripts/train_sd3_fast_pickscore.py:779`; the reference images of the adversarial loop are PNG files, README.md:114-128 of
the reference) returns for non-interlaced and Adam7-interlaced files: chunk walk (PNG 1.2 section 5), zlib inflate of the concatenated
IDAT data (Python's `zlib` is the pin for the library's own inflate), scan-line unfiltering (section 6: None / Sub / Up /
Average / Paeth, byte-wise modulo 256) and conversion to RGB (truecolour, truecolour + alpha: alpha dropped; greyscale:
replicated; palette: looked up).  Pinned to Pillow in tests/test_oracle_models.py::test_png_oracle_matches_pillow.
"""
import struct
import zlib

import numpy as np


class PngUnsupported(ValueError):
    pass


def parse(data):
    """-> dict(width, height, bit_depth, color_type, interlace, palette uint8 [n, 3] or None, idat bytes)."""
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise PngUnsupported("not a PNG file")
    pos, info, idat, palette = 8, None, [], None
    while pos + 8 <= len(data):
        ln, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + ln]
        pos += 12 + ln
        if typ == b"IHDR":
            w, h, bd, ct, cm, fm, il = struct.unpack(">IIBBBBB", body)
            info = dict(width=w, height=h, bit_depth=bd, color_type=ct, interlace=il)
        elif typ == b"PLTE":
            palette = np.frombuffer(body, dtype=np.uint8).reshape(-1, 3)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
    if info is None:
        raise PngUnsupported("no IHDR")
    info.update(palette=palette, idat=b"".join(idat))
    return info


CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}


def check_supported(info):
    bd, ct = info["bit_depth"], info["color_type"]
    if info["interlace"] not in (0, 1):
        raise PngUnsupported("unknown interlace method")
    if ct not in CHANNELS:
        raise PngUnsupported("bad colour type")
    ok = {0: (1, 2, 4, 8, 16), 2: (8, 16), 3: (1, 2, 4, 8), 4: (8, 16), 6: (8, 16)}[ct]
    if bd not in ok:
        raise PngUnsupported(f"bit depth {bd} with colour type {ct}")
    if ct == 3 and info["palette"] is None:
        raise PngUnsupported("palette image without PLTE")


def geometry(info):
    """-> (bytes per scan line, bytes per complete pixel as the filters see it: at least 1)"""
    bits = CHANNELS[info["color_type"]] * info["bit_depth"]
    return (info["width"] * bits + 7) // 8, max(1, bits // 8)


def unfilter(raw, height, rowbytes, bpp):
    """raw: uint8 [height * (1 + rowbytes)] filtered scan lines -> uint8 [height, rowbytes] (PNG 1.2 section 6.6)."""
    raw = np.frombuffer(raw, dtype=np.uint8).reshape(height, 1 + rowbytes)
    out = np.zeros((height, rowbytes), dtype=np.uint8)
    prior = np.zeros(rowbytes, dtype=np.int32)
    for y in range(height):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        rec = np.zeros(rowbytes, dtype=np.int32)
        if ft == 0:
            rec = line
        elif ft == 2:
            rec = (line + prior) & 255
        else:
            for i in range(rowbytes):
                a = rec[i - bpp] if i >= bpp else 0
                b = prior[i]
                c = prior[i - bpp] if i >= bpp else 0
                if ft == 1:
                    pr = a
                elif ft == 3:
                    pr = (a + b) >> 1
                elif ft == 4:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pr = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                else:
                    raise PngUnsupported(f"bad filter type {ft}")
                rec[i] = (line[i] + pr) & 255
        out[y] = rec
        prior = rec
    return out


def to_rgb(rows, info):
    """What Pillow's convert("RGB") yields: sub-byte samples are unpacked (MSB first) -- greyscale scaled to 0..255 (x 255 / 85 /
    17), palette indices looked up; 16-bit truecolour and 16-bit greyscale + alpha (raw mode LA;16B) keep the high byte of every
    sample, 16-bit greyscale (mode I;16) is CLIPPED to 255; alpha is dropped."""
    h, w, ct, bd = info["height"], info["width"], info["color_type"], info["bit_depth"]
    ch = CHANNELS[ct]
    if bd < 8:
        bits = np.unpackbits(rows, axis=1)[:, :w * bd].reshape(h, w, bd)
        px = (bits * (1 << np.arange(bd - 1, -1, -1))).sum(-1).astype(np.int32)[..., None]
        if ct == 0:
            px = px * (255 // ((1 << bd) - 1))
    elif bd == 16:
        s16 = rows.reshape(h, w, ch, 2).astype(np.int32)
        px = np.minimum(s16[..., 0] * 256 + s16[..., 1], 255) if ct == 0 else s16[..., 0]
    else:
        px = rows.reshape(h, w, ch).astype(np.int32)
    px = px.astype(np.uint8)
    if ct == 2:
        return px.copy()
    if ct == 6:
        return px[..., :3].copy()                        # Image.convert("RGB") drops the alpha channel
    if ct in (0, 4):
        return np.repeat(px[..., :1], 3, axis=2)
    pal = np.zeros((256, 3), dtype=np.uint8)
    pal[:len(info["palette"])] = info["palette"]
    return pal[px[..., 0]]


ADAM7 = ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2))   # x0, y0, dx, dy


def decode_rgb(data):
    """bytes of a PNG file -> uint8 [H, W, 3], equal to np.asarray(Image.open(...).convert("RGB"))."""
    info = parse(data)
    check_supported(info)
    raw = zlib.decompress(info["idat"])
    if not info["interlace"]:
        rowbytes, bpp = geometry(info)
        return to_rgb(unfilter(raw, info["height"], rowbytes, bpp), info)
    # Adam7 (PNG 1.2 section 8.2): seven reduced images, each with its own filtered scan lines, scattered into place
    out = np.zeros((info["height"], info["width"], 3), dtype=np.uint8)
    off = 0
    for x0, y0, dx, dy in ADAM7:
        pw, ph = -(-(info["width"] - x0) // dx), -(-(info["height"] - y0) // dy)
        if pw <= 0 or ph <= 0:
            continue
        sub = dict(info, width=pw, height=ph)
        rowbytes, bpp = geometry(sub)
        n = ph * (1 + rowbytes)
        out[y0::dy, x0::dx] = to_rgb(unfilter(raw[off:off + n], ph, rowbytes, bpp), sub)
        off += n
    return out
