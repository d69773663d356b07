"""Oracle (test infrastructure only): JPEG decode (sequential and progressive Huffman) restated in numpy, bit for bit what libjpeg(-turbo) -- i.e.
Pillow's `Image.open(path).convert("RGB")` at `scripts/train_sd3_fast_pickscore.py:779` -- produces with its default
settings: sequential Huffman entropy decoding (ITU T.81 F.2.2), dequantisation + the "islow" integer inverse DCT
(jidctint.c: 13-bit constants, 2 extra bits after the column pass), "fancy" triangle chroma upsampling for 4:2:0 / 4:2:2
(jdsample.c h2v2_fancy_upsample / h2v1_fancy_upsample), and the fixed-point YCbCr -> RGB tables of jdcolor.c.

libjpeg itself is a third-party dependency of Pillow that is not under /root/reference; the pin is Pillow's own decoder
on the same files (tests/test_oracle_models.py::test_jpeg_oracle_matches_pillow).  Pure-Python loops: small images only.
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7,
                   14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39,
                   46, 53, 60, 61, 54, 47, 55, 62, 63])


class JpegUnsupported(ValueError):
    pass


def parse(data):
    """Marker walk of a baseline (SOF0), extended-sequential (SOF1) or progressive (SOF2) 8-bit Huffman file.  Returns a dict
    with the frame geometry, the quantisation tables and, per scan, its component selectors, spectral / successive-
    approximation parameters, the Huffman tables and restart interval in force, and the offset of its entropy-coded data."""
    if data[:2] != b"\xff\xd8":
        raise JpegUnsupported("not a JPEG (no SOI)")
    pos, q, huff, frame, ri, adobe_transform, scans, progressive = 2, {}, {}, None, 0, None, [], False
    while pos < len(data):
        if data[pos] != 0xFF:
            pos += 1                                    # entropy-coded bytes of the previous scan: skip to the next marker
            continue
        m = data[pos + 1]
        if m == 0x00 or m == 0xFF or 0xD0 <= m <= 0xD7:  # stuffed byte / fill / RSTn inside entropy-coded data
            pos += 1 if m == 0xFF else 2
            continue
        pos += 2
        if m == 0x01:
            continue
        if m == 0xD9:                                   # EOI
            break
        ln = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + ln]
        if m == 0xDB:                                   # DQT
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                if pq:
                    raise JpegUnsupported("16-bit quantisation table")
                t = np.zeros(64, dtype=np.int32)
                t[ZIGZAG] = np.frombuffer(seg[i + 1:i + 65], dtype=np.uint8)
                q[tq] = t
                i += 65
        elif m in (0xC0, 0xC1, 0xC2):                   # SOF0 / SOF1 / SOF2
            if seg[0] != 8:
                raise JpegUnsupported("sample precision != 8")
            progressive = m == 0xC2
            h, w, n = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [dict(id=seg[6 + 3 * c], h=seg[7 + 3 * c] >> 4, v=seg[7 + 3 * c] & 15, tq=seg[8 + 3 * c]) for c in range(n)]
            frame = dict(height=h, width=w, comps=comps)
        elif m in (0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise JpegUnsupported("lossless / arithmetic / hierarchical JPEG")
        elif m == 0xC4:                                 # DHT
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                nsym = sum(counts)
                huff[(tc, th)] = (counts, list(seg[i + 17:i + 17 + nsym]))
                i += 17 + nsym
        elif m == 0xDD:
            ri = (seg[0] << 8) | seg[1]
        elif m == 0xEE and seg[:5] == b"Adobe":
            adobe_transform = seg[11]
        elif m == 0xDA:                                 # SOS
            ns = seg[0]
            sel = [dict(id=seg[1 + 2 * c], td=seg[2 + 2 * c] >> 4, ta=seg[2 + 2 * c] & 15) for c in range(ns)]
            ss, se, ah, al = seg[1 + 2 * ns], seg[2 + 2 * ns], seg[3 + 2 * ns] >> 4, seg[3 + 2 * ns] & 15
            if frame is None:
                raise JpegUnsupported("SOS before SOF")
            if not progressive and ns != len(frame["comps"]):
                raise JpegUnsupported("non-interleaved sequential file")
            scans.append(dict(sel=sel, ss=ss, se=se, ah=ah, al=al, huff=dict(huff), ri=ri, ecs=pos + ln))
            if not progressive:
                break
        pos += ln
    if not scans:
        raise JpegUnsupported("no SOS")
    return dict(frame=frame, q=q, huff=scans[0]["huff"], ri=scans[0]["ri"], scan=scans[0]["sel"], ecs=scans[0]["ecs"],
                adobe_transform=adobe_transform, progressive=progressive, scans=scans)


def _build_decoder(counts, symbols):
    """canonical Huffman code -> dict (length, code) -> symbol"""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


class _Bits:
    def __init__(self, data, pos):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def _fill(self):
        b = self.d[self.p] if self.p < len(self.d) else 0
        self.p += 1
        if b == 0xFF:
            nxt = self.d[self.p] if self.p < len(self.d) else 0
            if nxt == 0:
                self.p += 1                             # stuffed zero
            else:                                       # a marker: feed zeros (libjpeg does the same at the end of data)
                self.p -= 1
                b = 0
        self.acc = (self.acc << 8) | b
        self.n += 8

    def bit(self):
        if self.n == 0:
            self._fill()
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def restart(self):
        """byte-align and skip the RSTn marker"""
        self.acc, self.n = 0, 0
        while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _extend(v, t):
    return v - ((1 << t) - 1) if t and v < (1 << (t - 1)) else v


def entropy_decode(data, info):
    """-> list per component of int16 [blocks_h, blocks_w, 64] quantised coefficients in NATURAL order (padded to whole MCUs)."""
    fr = info["frame"]
    comps = fr["comps"]
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mcux, mcuy = -(-fr["width"] // (8 * hmax)), -(-fr["height"] // (8 * vmax))
    out = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), dtype=np.int16) for c in comps]
    if info.get("progressive"):
        return _entropy_decode_progressive(data, info, out, mcux, mcuy, hmax, vmax)
    dec = {k: _build_decoder(*v) for k, v in info["huff"].items()}
    sel = {s["id"]: s for s in info["scan"]}
    br = _Bits(data, info["ecs"])
    pred = [0] * len(comps)

    def sym(tbl):
        code = 0
        for length in range(1, 17):
            code = (code << 1) | br.bit()
            s = tbl.get((length, code))
            if s is not None:
                return s
        raise ValueError("bad Huffman code")

    n = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if info["ri"] and n and n % info["ri"] == 0:
                br.restart()
                pred = [0] * len(comps)
            n += 1
            for ci, c in enumerate(comps):
                s = sel[c["id"]]
                dct, act = dec[(0, s["td"])], dec[(1, s["ta"])]
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = out[ci][my * c["v"] + by, mx * c["h"] + bx]
                        t = sym(dct)
                        pred[ci] += _extend(br.bits(t), t)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = sym(act)
                            r, sz = rs >> 4, rs & 15
                            if sz == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(sz), sz)
                            k += 1
    return out


def _entropy_decode_progressive(data, info, out, mcux, mcuy, hmax, vmax):
    """ITU T.81 annex G (libjpeg jdphuff.c): DC first / refinement scans (possibly interleaved) and single-component AC
    first / refinement scans with end-of-band runs; coefficients accumulate over the scans in `out` (natural order)."""
    fr = info["frame"]
    comps = fr["comps"]
    idx_of = {c["id"]: i for i, c in enumerate(comps)}
    for sc in info["scans"]:
        dec = {k: _build_decoder(*v) for k, v in sc["huff"].items()}
        br = _Bits(data, sc["ecs"])
        ss, se, ah, al = sc["ss"], sc["se"], sc["ah"], sc["al"]
        cis = [idx_of[s["id"]] for s in sc["sel"]]

        def sym(tbl):
            code = 0
            for length in range(1, 17):
                code = (code << 1) | br.bit()
                s_ = tbl.get((length, code))
                if s_ is not None:
                    return s_
            raise ValueError("bad Huffman code")

        # the units of the scan: MCUs for an interleaved scan, the component's own blocks (ceil(w_c / 8) x ceil(h_c / 8))
        # for a single-component scan
        if len(cis) > 1:
            units = [(my, mx) for my in range(mcuy) for mx in range(mcux)]
        else:
            c = comps[cis[0]]
            wc = -(-(fr["width"] * c["h"]) // hmax)
            hc = -(-(fr["height"] * c["v"]) // vmax)
            units = [(by, bx) for by in range(-(-hc // 8)) for bx in range(-(-wc // 8))]
        pred = [0] * len(comps)
        eobrun = 0
        for n, (uy, ux) in enumerate(units):
            if sc["ri"] and n and n % sc["ri"] == 0:
                br.restart()
                pred = [0] * len(comps)
                eobrun = 0
            if len(cis) > 1:
                blocks = [(ci, uy * comps[ci]["v"] + by, ux * comps[ci]["h"] + bx) for ci in cis
                          for by in range(comps[ci]["v"]) for bx in range(comps[ci]["h"])]
            else:
                blocks = [(cis[0], uy, ux)]
            for ci, by, bx in blocks:
                blk = out[ci][by, bx]
                sel = next(s for s in sc["sel"] if idx_of[s["id"]] == ci)
                if ss == 0:                                       # DC scan
                    if ah == 0:
                        t = sym(dec[(0, sel["td"])])
                        pred[ci] += _extend(br.bits(t), t)
                        blk[0] = pred[ci] * (1 << al)
                    elif br.bit():
                        blk[0] |= (1 << al)
                    continue
                act = dec[(1, sel["ta"])]
                if ah == 0:                                       # AC first scan
                    if eobrun > 0:
                        eobrun -= 1
                        continue
                    k = ss
                    while k <= se:
                        rs = sym(act)
                        r, sz = rs >> 4, rs & 15
                        if sz:
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(sz), sz) * (1 << al)
                            k += 1
                        elif r == 15:
                            k += 16
                        else:
                            eobrun = (1 << r) + (br.bits(r) if r else 0) - 1
                            break
                    continue
                # AC refinement scan (jdphuff.c decode_mcu_AC_refine)
                p1, m1 = 1 << al, -(1 << al)
                k = ss
                if eobrun == 0:
                    while k <= se:
                        rs = sym(act)
                        r, sz = rs >> 4, rs & 15
                        val = 0
                        if sz:
                            val = p1 if br.bit() else m1          # size is always 1 here
                        elif r != 15:
                            eobrun = (1 << r) + (br.bits(r) if r else 0)
                            break
                        while k <= se:                             # skip r zero-history coefficients, refining the others
                            z = ZIGZAG[k]
                            if blk[z] != 0:
                                if br.bit() and (blk[z] & p1) == 0:
                                    blk[z] += p1 if blk[z] >= 0 else m1
                            else:
                                if r == 0:
                                    break
                                r -= 1
                            k += 1
                        if val and k <= se:
                            blk[ZIGZAG[k]] = val
                        k += 1
                if eobrun > 0:                                     # the rest of the band: only correction bits
                    while k <= se:
                        z = ZIGZAG[k]
                        if blk[z] != 0 and br.bit() and (blk[z] & p1) == 0:
                            blk[z] += p1 if blk[z] >= 0 else m1
                        k += 1
                    eobrun -= 1
    return out


# ---- jidctint.c (islow), CONST_BITS = 13, PASS1_BITS = 2 ----
_C = dict(f0_298631336=2446, f0_390180644=3196, f0_541196100=4433, f0_765366865=6270, f0_899976223=7373, f1_175875602=9633,
          f1_501321110=12299, f1_847759065=15137, f1_961570560=16069, f2_053119869=16819, f2_562915447=20995, f3_072711026=25172)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(d, shift_in, shift_out):
    """one pass of jpeg_idct_islow over the last axis of d (int64 [..., 8]); shift_in applies to the even part's DC terms"""
    z2, z3 = d[..., 2], d[..., 6]
    z1 = (z2 + z3) * _C["f0_541196100"]
    tmp2 = z1 + z3 * (-_C["f1_847759065"])
    tmp3 = z1 + z2 * _C["f0_765366865"]
    z2, z3 = d[..., 0], d[..., 4]
    tmp0 = (z2 + z3) << 13
    tmp1 = (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = d[..., 7], d[..., 5], d[..., 3], d[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * _C["f1_175875602"]
    tmp0 = tmp0 * _C["f0_298631336"]
    tmp1 = tmp1 * _C["f2_053119869"]
    tmp2 = tmp2 * _C["f3_072711026"]
    tmp3 = tmp3 * _C["f1_501321110"]
    z1 = z1 * (-_C["f0_899976223"])
    z2 = z2 * (-_C["f2_562915447"])
    z3 = z3 * (-_C["f1_961570560"]) + z5
    z4 = z4 * (-_C["f0_390180644"]) + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    o = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2,
                  tmp10 - tmp3], axis=-1)
    return _descale(o, shift_out)


def idct_islow(coefs, qt):
    """coefs int16 [..., 64] (natural order), qt int32 [64] -> uint8 [..., 8, 8] samples (level shift + clamp)."""
    d = coefs.astype(np.int64) * qt.astype(np.int64)
    d = d.reshape(d.shape[:-1] + (8, 8))                       # [row, col]
    ws = _idct_1d(np.swapaxes(d, -1, -2), 0, 13 - 2)            # pass 1: columns (each column is a length-8 vector)
    ws = np.swapaxes(ws, -1, -2)
    o = _idct_1d(ws, 0, 13 + 2 + 3)                             # pass 2: rows
    return np.clip(o + 128, 0, 255).astype(np.uint8)


def _planes(info, coefs):
    fr = info["frame"]
    planes = []
    for c, co in zip(fr["comps"], coefs):
        px = idct_islow(co, info["q"][c["tq"]])                 # [bh, bw, 8, 8]
        bh, bw = px.shape[:2]
        planes.append(px.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8))
    return planes


def _h2v1_fancy(p, w_down):
    """jdsample.c h2v1_fancy_upsample on rows of width w_down -> 2 w_down"""
    p = p[:, :w_down].astype(np.int32)
    if w_down <= 2:                                       # jinit_upsampler: fancy only when downsampled_width > 2
        return np.repeat(p, 2, axis=1)
    out = np.zeros((p.shape[0], 2 * w_down), dtype=np.int32)
    left = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    out[:, 0::2] = (3 * p + left + 1) >> 2
    out[:, 1::2] = (3 * p + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out


def _h2v2_fancy(p, w_down, h_down):
    """jdsample.c h2v2_fancy_upsample: vertical 3:1 blend with the nearer neighbouring row (edge rows replicated by the
    main controller's context handling), then the horizontal triangle with the 8 / 7 rounding pair."""
    p = p[:h_down, :w_down].astype(np.int32)
    if w_down <= 2:                                       # jinit_upsampler: plain 2x2 replication for very narrow planes
        return np.repeat(np.repeat(p, 2, axis=0), 2, axis=1)
    up = np.concatenate([p[:1], p[:-1]], axis=0)
    dn = np.concatenate([p[1:], p[-1:]], axis=0)
    out = np.zeros((2 * h_down, 2 * w_down), dtype=np.int32)
    for v, other in ((0, up), (1, dn)):
        s = 3 * p + other                                        # "colsum" of every column
        o = np.zeros((h_down, 2 * w_down), dtype=np.int32)
        last = np.concatenate([s[:, :1], s[:, :-1]], axis=1)
        nxt = np.concatenate([s[:, 1:], s[:, -1:]], axis=1)
        o[:, 0::2] = (3 * s + last + 8) >> 4
        o[:, 1::2] = (3 * s + nxt + 7) >> 4
        o[:, 0] = (s[:, 0] * 4 + 8) >> 4
        o[:, -1] = (s[:, -1] * 4 + 7) >> 4
        out[v::2] = o
    return out


def _fix(x):
    return int(x * 65536 + 0.5)


def ycc_to_rgb(y, cb, cr):
    """jdcolor.c build_ycc_rgb_table + ycc_rgb_convert"""
    y, cb, cr = y.astype(np.int32), cb.astype(np.int32) - 128, cr.astype(np.int32) - 128
    r = y + ((_fix(1.40200) * cr + 32768) >> 16)
    b = y + ((_fix(1.77200) * cb + 32768) >> 16)
    g = y + ((-_fix(0.34414) * cb + 32768 - _fix(0.71414) * cr) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def assemble_rgb(info, coefs):
    """Back end shared by `decode_rgb` and the tests of the library's host entropy decoder: per-component coefficient
    blocks -> uint8 [H, W, 3]."""
    fr = info["frame"]
    H, W, comps = fr["height"], fr["width"], fr["comps"]
    planes = _planes(info, coefs)
    if len(comps) == 1:
        g = planes[0][:H, :W]
        return np.stack([g, g, g], axis=-1)
    if len(comps) != 3 or info["adobe_transform"] == 0:
        raise JpegUnsupported("only grayscale and YCbCr files")
    hs, vs = [c["h"] for c in comps], [c["v"] for c in comps]
    if (hs[1], vs[1]) != (hs[2], vs[2]) or hs[1] != 1 or vs[1] != 1:
        raise JpegUnsupported("chroma sampling factors other than 1x1")
    y = planes[0][:H, :W]
    if (hs[0], vs[0]) == (1, 1):
        cb, cr = planes[1][:H, :W], planes[2][:H, :W]
    elif (hs[0], vs[0]) == (2, 1):
        wd = -(-W // 2)
        cb, cr = (_h2v1_fancy(p[:H], wd)[:, :W] for p in planes[1:])
    elif (hs[0], vs[0]) == (2, 2):
        wd, hd = -(-W // 2), -(-H // 2)
        cb, cr = (_h2v2_fancy(p, wd, hd)[:H, :W] for p in planes[1:])
    else:
        raise JpegUnsupported(f"luma sampling {hs[0]}x{vs[0]}")
    return ycc_to_rgb(y, cb, cr)


def decode_rgb(data):
    """bytes of a sequential or progressive Huffman JPEG -> uint8 [H, W, 3], equal to
    np.asarray(Image.open(...).convert("RGB"))."""
    info = parse(data)
    if len(info["frame"]["comps"]) not in (1, 3):
        raise JpegUnsupported("component count other than 1 or 3")
    return assemble_rgb(info, entropy_decode(data, info))
