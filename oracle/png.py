"""Oracle (test infrastructure only): PNG decode restated in numpy -- what Pillow's `Image.open(path).convert("RGB")`
(`scripts/train_sd3_fast_pickscore.py:779`; the reference images of the adversarial loop are PNG files, README.md:114-128 of
the reference) returns for non-interlaced 8-bit files: chunk walk (PNG 1.2 section 5), zlib inflate of the concatenated
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
    if info["interlace"]:
        raise PngUnsupported("Adam7-interlaced PNG")
    if info["bit_depth"] != 8:
        raise PngUnsupported("bit depth other than 8")
    if info["color_type"] not in CHANNELS:
        raise PngUnsupported("bad colour type")
    if info["color_type"] == 3 and info["palette"] is None:
        raise PngUnsupported("palette image without PLTE")


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
    h, w, ct = info["height"], info["width"], info["color_type"]
    px = rows.reshape(h, w, CHANNELS[ct])
    if ct == 2:
        return px.copy()
    if ct == 6:
        return px[..., :3].copy()                        # Image.convert("RGB") drops the alpha channel
    if ct == 0:
        return np.repeat(px, 3, axis=2)
    if ct == 4:
        return np.repeat(px[..., :1], 3, axis=2)
    pal = np.zeros((256, 3), dtype=np.uint8)
    pal[:len(info["palette"])] = info["palette"]
    return pal[px[..., 0]]


def decode_rgb(data):
    """bytes of a PNG file -> uint8 [H, W, 3], equal to np.asarray(Image.open(...).convert("RGB"))."""
    info = parse(data)
    check_supported(info)
    ch = CHANNELS[info["color_type"]]
    rowbytes = info["width"] * ch
    raw = zlib.decompress(info["idat"])
    return to_rgb(unfilter(raw, info["height"], rowbytes, ch), info)
