"""Reference-image path timings: Pillow (decode + resize + ToTensor on the host, then upload) vs the hybrid decoders of this
repo (host entropy decode, device back end, device resize), 1024 x 1024 and 2048 x 2048 PNG / JPEG -> 512 x 512 float tensor."""
import io
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from PIL import Image  # noqa: E402

from adv_grpo_b200 import jpeg, ops, png  # noqa: E402
from jpeg_util import _jpeg_bytes  # noqa: E402
from png_util import pillow_png  # noqa: E402

DEV = "cuda"


def pil_path(data):
    img = Image.open(io.BytesIO(data)).convert("RGB").resize((512, 512), Image.BILINEAR)
    return torch.from_numpy(np.asarray(img).copy()).permute(2, 0, 1).float().div_(255.0).to(DEV)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


def dev_ms(fn, n=5):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for size in (1024, 2048):
    p, j = pillow_png(size, size, "RGB", seed=size), _jpeg_bytes(size, size, seed=size, quality=90, subsampling=2)
    for name, data, dec in (("png", p, png.decode_png_to_device), ("jpeg", j, jpeg.decode_jpeg_to_device)):
        ours = lambda: ops.pil_resize_bilinear(dec(data, DEV), 512, 512)
        a, b = pil_path(data), ours()
        assert torch.equal(a, b), name
        raw = dec(data, DEV)
        print(f"{name} {size}x{size} ({len(data) / 1e6:.2f} MB): Pillow path {timed(lambda: pil_path(data)):.1f} ms, hybrid path "
              f"{timed(ours):.1f} ms wall (device resize alone {dev_ms(lambda: ops.pil_resize_bilinear(raw, 512, 512)):.3f} ms)")

# one prompt = 8 reference files (mini_num_image_per_prompt): host stages one after the other vs on the thread pool
import json  # noqa: E402
import tempfile  # noqa: E402

from adv_grpo_b200.reference_images import ReferenceImageIndex  # noqa: E402

tmp = tempfile.mkdtemp()
names = []
for i in range(8):
    names.append(f"r{i}.png")
    with open(os.path.join(tmp, names[-1]), "wb") as f:
        f.write(pillow_png(1024, 1024, "RGB", seed=100 + i))
with open(os.path.join(tmp, "index.json"), "w") as f:
    json.dump({"p": names}, f)
for label, kw in (("Pillow on the host, one file after the other (reference)", dict(device="cpu", host_threads=1)),
                  ("Pillow on the host, 8 threads", dict(device="cpu")), ("hybrid, host stages serial", dict(device=DEV, host_threads=1)),
                  ("hybrid, host stages on 8 threads", dict(device=DEV))):
    idx = ReferenceImageIndex(os.path.join(tmp, "index.json"), tmp, size=512, cache=False, **kw)
    fn = (lambda: idx("p").to(DEV)) if kw["device"] == "cpu" else (lambda: idx("p"))
    print(f"prompt with 8 x 1024x1024 PNG -> [8, 3, 512, 512] on the device: {label} {timed(fn, n=3):.1f} ms ({os.cpu_count()} host cores)")
