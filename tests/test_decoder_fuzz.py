"""Memory safety of the host halves of the image decoders (they parse untrusted file bytes): an AddressSanitizer build of
`csrc/jpeg.cu` + `csrc/png.cu` is fed mutated PNG / JPEG files through the C-ABI with exact-size buffers
(`tests/fuzz/fuzz_image_decoders.cpp`), and the corrupt headers that fuzzing found are kept as explicit cases.  CPU only."""
import os
import struct
import subprocess
import sys
import zlib

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _seed_files(folder):
    from jpeg_util import _jpeg_bytes
    from png_util import handmade_png, pillow_png
    files = {f"p{i}.png": handmade_png(13, 17, ct, seed=ct, bd=bd, interlace=il, level=lv)[0]
             for i, (ct, bd, il, lv) in enumerate(((2, 8, 0, 6), (6, 8, 1, 9), (3, 4, 0, 1), (0, 16, 1, 6), (4, 8, 0, 0), (3, 1, 1, 6),
                                                   (4, 16, 0, 6)))}
    files["pp.png"] = pillow_png(40, 56, "RGB", seed=1)
    for i, kw in enumerate((dict(quality=80), dict(quality=60, progressive=True), dict(subsampling=0), dict(gray=True),
                            dict(subsampling=1, progressive=True), dict(quality=95, subsampling=2, progressive=True),
                            dict(gray=True, progressive=True))):
        files[f"j{i}.jpg"] = _jpeg_bytes(40 + 3 * i, 48 - 5 * i, seed=i, **kw)
    paths = []
    for name, data in files.items():
        paths.append(os.path.join(folder, name))
        with open(paths[-1], "wb") as f:
            f.write(data)
    return paths


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")
def test_host_decoders_survive_mutated_files_under_asan(tmp_path):
    csrc = os.path.join(ROOT, "adv_grpo_b200", "csrc")
    exe = str(tmp_path / "fuzz")
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-g", "-std=c++17", "--expt-relaxed-constexpr",
           "-Xcompiler", "-fsanitize=address,-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "include"), "-o", exe,
           os.path.join(HERE, "fuzz", "fuzz_image_decoders.cpp")]
    cmd += [os.path.join(csrc, f) for f in ("core.cu", "jpeg.cu", "png.cu")] + ["-lasan"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr + r.stdout).lower():
        pytest.skip("AddressSanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    seeds = _seed_files(str(tmp_path))
    env = dict(os.environ, ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0")
    probe = subprocess.run([exe, "0", seeds[0]], capture_output=True, text=True, env=env, timeout=120)
    if probe.returncode != 0 and "decoded" not in probe.stdout and "does not decode" not in probe.stderr:
        pytest.skip("the AddressSanitizer runtime cannot start in this sandbox: " + probe.stderr[-200:])
    r = subprocess.run([exe, "1500"] + seeds, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "decoded" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])
    decoded, rejected = (int(x) for x in r.stdout.split()[1::2])
    assert decoded > 1000 and rejected > 1000          # the mutations reach both the decode loops and the error paths


def test_corrupt_headers_are_rejected():
    from adv_grpo_b200 import _lib, jpeg as jpeg_b, png as png_b
    from jpeg_util import _jpeg_bytes
    from png_util import _chunk, handmade_png
    good = _jpeg_bytes(32, 32, seed=0, quality=80)
    assert jpeg_b.entropy_decode(good)[0] is not None
    bad = []
    # DHT with three codes of length 1: not a prefix code (this once overflowed the 9-bit lookahead table)
    b = bytearray(good)
    o = b.index(b"\xff\xc4") + 5                      # counts of the codes of length 1, 2, 3 ...: 0 1 5 1 ... in the standard DC table
    assert b[o] == 0 and b[o + 2] == 5
    b[o], b[o + 2] = 3, 2                              # same number of symbols, three codes of length 1
    bad.append(bytes(b))
    # a second SOF naming one component after the first one named three
    i = good.index(b"\xff\xc0")
    ln = struct.unpack(">H", good[i + 2:i + 4])[0]
    sof = bytearray(good[i:i + 2 + ln])
    sof[2:4] = struct.pack(">H", 11)
    sof[9] = 1
    sos = good.index(b"\xff\xda")
    bad.append(good[:sos] + bytes(sof[:13]) + good[sos:])
    # 4.3 G pixels: beyond Pillow's decompression-bomb bound
    b = bytearray(good)
    b[i + 5:i + 9] = struct.pack(">HH", 65535, 65535)
    bad.append(bytes(b))
    for data, why in zip(bad, ("not a prefix code", "more than one SOF", "larger than")):
        with pytest.raises(_lib.AdvGrpoError, match=why):
            jpeg_b.entropy_decode(data)
    # PNG: a scan line whose filter type is 5; an image beyond the size bound
    w, h = 4, 3
    raw = bytearray(b"".join(bytes([0]) + bytes(3 * w) for _ in range(h)))
    raw[1 + 3 * w] = 5
    head = b"\x89PNG\r\n\x1a\n"
    png = (head + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + _chunk(b"IDAT", zlib.compress(bytes(raw))) +
           _chunk(b"IEND", b""))
    with pytest.raises(_lib.AdvGrpoError, match="filter type"):
        png_b.inflate(png)
    huge = head + _chunk(b"IHDR", struct.pack(">IIBBBBB", 30000, 30000, 8, 2, 0, 0, 0)) + _chunk(b"IDAT", b"x") + _chunk(b"IEND", b"")
    with pytest.raises(_lib.AdvGrpoError, match="larger than"):
        png_b.png_info(huge)
    assert png_b.inflate(handmade_png(5, 5, 2)[0])[0] is not None


def _mutate(rng, data):
    b = bytearray(data)
    k = int(rng.integers(0, 4))
    if k == 0:
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
    elif k == 1:
        b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
    elif k == 2:
        i = int(rng.integers(0, len(b)))
        del b[i:int(rng.integers(i, min(len(b), i + 20)))]
    else:
        i, j, n = int(rng.integers(0, len(b))), int(rng.integers(0, len(b))), int(rng.integers(1, 16))
        b[i:i + n] = b[j:j + n]
    return bytes(b)


def _pillow_rgb(data):
    import io
    import warnings
    import numpy as np
    from PIL import Image
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        except Exception:
            return None


def test_accepted_jpeg_mutants_equal_pillow():
    """Differential fuzz: whatever corrupt file the host entropy decoder TAKES must come out as Pillow (libjpeg-turbo) decodes it;
    everything it refuses goes to Pillow anyway.  This is what the strict checks of `parse_jpeg` / the decoders are for: scan
    segments that do not end on their last block, restart markers out of place, Huffman tables libjpeg refuses, unknown
    markers, incomplete progressive scripts (libjpeg would smooth), coefficients beyond the range of 8-bit samples (the
    16-bit SIMD inverse DCT saturates where 32-bit arithmetic does not).  Pixel stage: the numpy oracle's back end."""
    import numpy as np
    from adv_grpo_b200 import _lib, jpeg as jpeg_b
    from jpeg_util import _jpeg_bytes
    from oracle import jpeg as jpeg_o
    kws = (dict(quality=80), dict(quality=60, progressive=True), dict(subsampling=0), dict(gray=True), dict(subsampling=1, progressive=True),
           dict(quality=95, subsampling=2, progressive=True), dict(quality=70, subsampling=1), dict(gray=True, progressive=True))
    seeds = [_jpeg_bytes(24 + 3 * i, 40 - 3 * i, seed=i, **kw) for i, kw in enumerate(kws)]
    rng = np.random.default_rng(2024)
    taken = 0
    for it in range(800):
        data = _mutate(rng, seeds[it % len(seeds)])
        try:
            coefs, _, _ = jpeg_b.coefficients_as_numpy(data)
        except _lib.AdvGrpoError:
            continue
        if coefs is None:
            continue
        taken += 1
        got = jpeg_o.assemble_rgb(jpeg_o.parse(data), coefs)
        ref = _pillow_rgb(data)
        assert ref is not None and ref.shape == got.shape and np.array_equal(ref, got), f"mutant {it} of seed {it % len(seeds)}"
    assert taken > 100


def test_accepted_png_mutants_equal_pillow():
    """The same for PNG, with the chunk checksums of every mutant repaired (the decoder verifies them all, Pillow only those in
    front of the image data): zlib-level damage (Huffman tables zlib refuses, Adler-32, stream length), bad filter types, chunk
    order.  Pixel stage: the numpy oracle on the library's inflate output."""
    import numpy as np
    from adv_grpo_b200 import _lib, png as png_b
    from oracle import png as png_o
    from png_util import handmade_png
    seeds = [handmade_png(9, 11, ct, seed=ct, bd=bd, interlace=il, level=lv, split=sp)[0]
             for ct, bd, il, lv, sp in ((2, 8, 0, 6, None), (6, 8, 1, 9, 20), (3, 4, 0, 1, None), (0, 16, 1, 6, 7), (4, 8, 0, 0, None),
                                        (3, 1, 1, 6, None), (4, 16, 0, 6, 50), (3, 8, 0, 6, None))]

    def repair(b):
        b, pos = bytearray(b), 8
        while pos + 12 <= len(b):
            ln = struct.unpack(">I", b[pos:pos + 4])[0]
            if pos + 12 + ln > len(b):
                break
            b[pos + 8 + ln:pos + 12 + ln] = struct.pack(">I", zlib.crc32(bytes(b[pos + 4:pos + 8 + ln])))
            pos += 12 + ln
        return bytes(b)

    def pixels(raw, pal, info):
        inf = dict(width=info.width, height=info.height, bit_depth=info.bit_depth, color_type=info.color_type, interlace=info.interlace,
                   palette=pal.numpy().reshape(256, 3))
        rawb, out, off = raw.numpy().tobytes(), np.zeros((info.height, info.width, 3), np.uint8), 0
        for x0, y0, dx, dy in (png_o.ADAM7 if info.interlace else ((0, 0, 1, 1),)):
            pw, ph = -(-(info.width - x0) // dx), -(-(info.height - y0) // dy)
            if pw <= 0 or ph <= 0:
                continue
            sub = dict(inf, width=pw, height=ph)
            rb, bpp = png_o.geometry(sub)
            out[y0::dy, x0::dx] = png_o.to_rgb(png_o.unfilter(rawb[off:off + ph * (1 + rb)], ph, rb, bpp), sub)
            off += ph * (1 + rb)
        return out

    rng = np.random.default_rng(7)
    taken = 0
    for it in range(1600):
        data = repair(_mutate(rng, seeds[it % len(seeds)]))
        try:
            raw, pal, info = png_b.inflate(data)
        except _lib.AdvGrpoError:
            continue
        if raw is None:
            continue
        taken += 1
        ref = _pillow_rgb(data)
        got = pixels(raw, pal, info)
        assert ref is not None and ref.shape == got.shape and np.array_equal(ref, got), f"mutant {it} of seed {it % len(seeds)}"
    assert taken > 100


def test_host_decoders_are_thread_safe():
    """`ReferenceImageIndex` runs the host stages of a prompt's files on a thread pool: the same files decoded from eight
    threads at once give the bytes a serial pass gives (per-thread error string, no shared state in the decoders)."""
    from concurrent.futures import ThreadPoolExecutor
    from adv_grpo_b200 import _lib, jpeg as jpeg_b, png as png_b
    from jpeg_util import _jpeg_bytes
    from png_util import handmade_png
    jobs = [("p", handmade_png(150 + 7 * i, 200 - 5 * i, (2, 6, 0, 3)[i % 4], seed=i, interlace=i % 2)[0]) for i in range(8)]
    jobs += [("j", _jpeg_bytes(160 + 8 * i, 200 - 8 * i, seed=i, quality=85, subsampling=i % 3, progressive=bool(i % 2))) for i in range(8)]
    jobs += [("p", b"\x89PNG\r\n\x1a\n" + b"garbage" * 9), ("j", b"\xff\xd8" + b"garbage" * 9)]

    def run(job):
        kind, data = job
        try:
            out = png_b.inflate(data) if kind == "p" else jpeg_b.entropy_decode(data)
        except _lib.AdvGrpoError as e:
            return str(e)
        return out[0].numpy().tobytes()

    serial = [run(j) for j in jobs]
    with ThreadPoolExecutor(max_workers=8) as ex:
        for _ in range(3):
            assert list(ex.map(run, jobs)) == serial
    assert "png" in serial[-2] and "jpeg" in serial[-1]


def test_host_decoder_on_the_gpu_suite_files():
    """The files of `test_kernels_gpu.py::test_jpeg_decode_bit_exact_with_pillow` (restart intervals, progressive scripts,
    narrow planes, 1 x 1), host half only: the library's coefficients pushed through the numpy oracle's back end equal Pillow,
    so a change of the host decoder is checked here without a GPU."""
    import numpy as np
    from adv_grpo_b200 import jpeg as jpeg_b
    from jpeg_util import _jpeg_bytes
    from oracle import jpeg as jpeg_o
    cases = [(512, 512, dict(quality=90, subsampling=2)), (333, 517, dict(quality=75, subsampling=1)), (600, 401, dict(quality=95, subsampling=0)),
             (480, 640, dict(quality=60, subsampling=2, restart_marker_blocks=4)), (257, 129, dict(quality=80, gray=True)),
             (1, 1, dict(quality=90, subsampling=2)), (17, 9, dict(quality=100, subsampling=2)),
             (512, 768, dict(quality=85, subsampling=2, progressive=True)), (301, 203, dict(quality=92, subsampling=1, progressive=True)),
             (480, 640, dict(quality=70, subsampling=0, progressive=True, restart_marker_blocks=8)),
             (47, 3, dict(quality=100, subsampling=2)), (16, 4, dict(quality=90, subsampling=1)),
             (30, 2, dict(quality=80, subsampling=2, progressive=True)), (9, 5, dict(quality=90, subsampling=2))]
    for h, w, kw in cases:
        data = _jpeg_bytes(h, w, seed=h + w, **kw)
        coefs, _, info = jpeg_b.coefficients_as_numpy(data)
        assert coefs is not None and (info.height, info.width) == (h, w), (h, w, kw)
        assert np.array_equal(jpeg_o.assemble_rgb(jpeg_o.parse(data), coefs), _pillow_rgb(data)), (h, w, kw)
