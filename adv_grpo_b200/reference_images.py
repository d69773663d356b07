"""Reference ("real") images of the adversarial loop: the `{prompt: [file, ...]}` index JSON (`config.json_path`,
README.md:114-128 of the reference) and the per-batch loading of `scripts/train_sd3_fast_pickscore.py:705-707,
773-799`: `Image.open(root / name).convert("RGB")` -> `transforms.Resize((512, 512))` (PIL bilinear with Pillow's
antialiasing) -> `ToTensor()` ([0,1] float32, CHW) -> stacked on the device.  With a CUDA device: PNG files (what the
reference's reference-image generator writes) are inflated by the library's own host inflate and unfiltered + converted to
RGB on the GPU (`png.decode_png_to_device`); JPEG files (sequential or progressive) are entropy-decoded by the library's own
host Huffman decoder and everything after that (inverse DCT, chroma upsampling, colour conversion:
`jpeg.decode_jpeg_to_device`) runs on the GPU -- both byte-exact with Pillow; files outside those decoders' subsets
(arithmetic-coded or CMYK JPEG, other formats) are decoded by Pillow on the host and their bytes
uploaded; the antialiased bilinear resize +
ToTensor always run on the GPU (`ops.pil_resize_bilinear`, Pillow bit-exact).  Results are cached per prompt because
the files never change, which the reference does not do (it re-opens every file of the prompt for every batch), and
the host stages of a prompt's files run in parallel on a thread pool.  SURVEY.md section 8f rank 3."""
import json
import os

import numpy as np
import torch


class ReferenceImageIndex:
    def __init__(self, json_path, image_root, size=512, device="cuda", default_image=None, cache=True, host_threads=None):
        with open(json_path, "r", encoding="utf-8") as f:
            self.index = json.load(f)                                  # train_pick:705-707
        self.root, self.size, self.device = image_root, int(size), device
        self.default_image = default_image                            # the reference hard-codes a fallback file (:784)
        self._cache = {} if cache else None
        self.host_threads = min(8, os.cpu_count() or 1) if host_threads is None else int(host_threads)
        self._pool = None

    def __contains__(self, prompt):
        return prompt in self.index

    def _host_stage(self, path):
        """Everything that needs no GPU, safe to run on a worker thread (the C calls and Pillow release the GIL): read the
        file and run the library's host inflate / Huffman decode, or Pillow for what they do not take.
        -> ("png", raw, palette, info) | ("jpeg", coefs, qtabs, info) | ("rgb", uint8 ndarray [H, W, 3])"""
        from PIL import Image
        if torch.device(self.device).type == "cuda":
            from . import _lib, jpeg, png
            try:
                with open(path, "rb") as f:
                    data = f.read()
                if data[:8] == b"\x89PNG\r\n\x1a\n":                       # the reference images of the loop are PNG files
                    raw, pal, info = png.inflate(data)
                    if raw is not None:
                        return ("png", raw, pal, info)
                elif data[:2] == b"\xff\xd8":
                    coefs, qt, info = jpeg.entropy_decode(data)
                    if coefs is not None:
                        return ("jpeg", coefs, qt, info)
            except (OSError, _lib.AdvGrpoError):
                pass                                                   # unreadable / corrupt / outside the subset: Pillow decides
        try:
            img = Image.open(path).convert("RGB")
        except Exception as e:                                         # train_pick:781-786: fall back to the default image
            if self.default_image is None:
                raise FileNotFoundError(f"reference image {path} could not be opened ({e}) and no default_image is set")
            img = Image.open(self.default_image).convert("RGB")
        return ("rgb", np.asarray(img, dtype=np.uint8))

    def _device_stage(self, item):
        from PIL import Image
        if torch.device(self.device).type == "cuda":
            from . import jpeg, ops, png
            if item[0] == "png":                                       # wavefront unfiltering + RGB conversion on the GPU
                raw = png.unfilter_on_device(item[1], item[2], item[3], self.device)
            elif item[0] == "jpeg":                                    # inverse DCT, chroma upsampling, colour conversion on the GPU
                raw = jpeg.idct_on_device(item[1], item[2], item[3], self.device)
            else:                                                      # decoded by Pillow: bytes -> device
                raw = torch.from_numpy(item[1].copy()).to(self.device)
            return ops.pil_resize_bilinear(raw, self.size, self.size)  # resize + ToTensor on the GPU
        img = Image.fromarray(item[1]).resize((self.size, self.size), Image.BILINEAR)   # CPU device (host tests): Resize on PIL
        arr = np.asarray(img, dtype=np.uint8)                          # HWC
        return torch.from_numpy(arr.copy()).permute(2, 0, 1).float().div_(255.0)   # ToTensor()

    def _load(self, path):
        return self._device_stage(self._host_stage(path))

    def _load_many(self, paths):
        """The host stages of a prompt's files run side by side on a small thread pool (entropy decoding is serial per file
        and is most of the time: 8 files of 1024 x 1024 take 8 x 12 ms one after the other); the device stages follow in
        order on the caller's thread and stream."""
        if len(paths) < 2 or self.host_threads < 2:
            return [self._load(p) for p in paths]
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.host_threads, thread_name_prefix="refimg")
        futures = [self._pool.submit(self._host_stage, p) for p in paths]
        return [self._device_stage(f.result()) for f in futures]

    def __call__(self, prompt, n=None):
        """-> float32 [len(files) or n, 3, size, size] in [0, 1] on the device (train_pick:797-799).  With `n`, the
        file list is cycled / truncated to n images (the reference assumes len(files) == mini_num_image_per_prompt)."""
        if prompt not in self.index:
            raise KeyError(f"no reference images for prompt {prompt!r} in the index (train_pick:773,787-789)")
        if self._cache is not None and prompt in self._cache:
            t = self._cache[prompt]
        else:
            files = self.index[prompt]
            t = torch.stack(self._load_many([os.path.join(self.root, f) for f in files])).to(self.device, torch.float32)
            if self._cache is not None:
                self._cache[prompt] = t
        if n is not None and t.shape[0] != n:
            t = t[torch.arange(n, device=t.device) % t.shape[0]]
        return t

    def as_trainer_fn(self, prompts):
        """`GRPOTrainer(reference_image_fn=...)` hook: (prompt_index, n, size) -> images."""
        def fn(prompt_index, n, size):
            assert size == self.size, f"index built for {self.size}px, trainer asked for {size}px"
            return self(prompts[int(prompt_index)], n)
        return fn
