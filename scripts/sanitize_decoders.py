"""Small end-to-end pass over the image-decode kernels for compute-sanitizer (memcheck / racecheck): PNG wavefront
unfiltering (two bands, every filter type, Adam7, sub-byte and 16-bit depths), the JPEG back end (4:4:4 / 4:2:2 / 4:2:0 /
grey, sequential and progressive) and the Pillow-exact resize; every result is compared with Pillow.

    compute-sanitizer --tool memcheck  python scripts/sanitize_decoders.py
    compute-sanitizer --tool racecheck python scripts/sanitize_decoders.py
"""
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from PIL import Image  # noqa: E402

from adv_grpo_b200 import jpeg, ops, png  # noqa: E402
from jpeg_util import _jpeg_bytes  # noqa: E402
from png_util import handmade_png  # noqa: E402


def check(name, data, dec):
    ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    got = dec(data, "cuda")
    assert got is not None and torch.equal(got.cpu(), torch.from_numpy(ref.copy())), name
    small = ops.pil_resize_bilinear(got, 64, 48)
    pil = np.asarray(Image.fromarray(ref).resize((48, 64), Image.BILINEAR))            # PIL size = (width, height)
    assert torch.equal((small * 255).round().to(torch.uint8).cpu(), torch.from_numpy(pil.copy()).permute(2, 0, 1)), name + " resize"
    print("ok", name, tuple(ref.shape))


for ct, bd, il, h, w in ((2, 8, 0, 1100, 37), (6, 8, 1, 130, 61), (0, 1, 0, 70, 83), (3, 4, 1, 40, 40), (2, 16, 0, 33, 20), (4, 16, 1, 19, 23)):
    check(f"png ct{ct} bd{bd} il{il}", handmade_png(h, w, ct, seed=ct + bd, bd=bd, interlace=il)[0], png.decode_png_to_device)
for kw in (dict(subsampling=0), dict(subsampling=1), dict(subsampling=2), dict(gray=True), dict(subsampling=2, progressive=True)):
    check(f"jpeg {kw}", _jpeg_bytes(75, 131, seed=3, quality=85, **kw), jpeg.decode_jpeg_to_device)
torch.cuda.synchronize()
print("done")
