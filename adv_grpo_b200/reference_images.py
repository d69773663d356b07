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
the files never change, which the reference does not do (it re-opens every file of the prompt for every batch).
SURVEY.md section 8f rank 3."""
import json
import os

import numpy as np
import torch


class ReferenceImageIndex:
    def __init__(self, json_path, image_root, size=512, device="cuda", default_image=None, cache=True):
        with open(json_path, "r", encoding="utf-8") as f:
            self.index = json.load(f)                                  # train_pick:705-707
        self.root, self.size, self.device = image_root, int(size), device
        self.default_image = default_image                            # the reference hard-codes a fallback file (:784)
        self._cache = {} if cache else None

    def __contains__(self, prompt):
        return prompt in self.index

    def _load(self, path):
        from PIL import Image
        cuda = torch.device(self.device).type == "cuda"
        if cuda:                                                       # baseline JPEG: host Huffman decode, the rest on the GPU
            from . import _lib, jpeg, ops, png
            raw = None
            try:
                with open(path, "rb") as f:
                    data = f.read()
                if data[:8] == b"\x89PNG\r\n\x1a\n":                       # the reference images of the loop are PNG files
                    raw = png.decode_png_to_device(data, self.device)     # None: outside the subset -> Pillow below
                elif data[:2] == b"\xff\xd8":
                    raw = jpeg.decode_jpeg_to_device(data, self.device)   # None: CMYK / arithmetic / ... -> Pillow below
            except (OSError, _lib.AdvGrpoError):
                raw = None                                             # unreadable / corrupt: Pillow decides (and falls back)
            if raw is not None:
                return ops.pil_resize_bilinear(raw, self.size, self.size)
        try:
            img = Image.open(path).convert("RGB")
        except Exception as e:                                         # train_pick:781-786: fall back to the default image
            if self.default_image is None:
                raise FileNotFoundError(f"reference image {path} could not be opened ({e}) and no default_image is set")
            img = Image.open(self.default_image).convert("RGB")
        if cuda:                                                       # decoded bytes -> device, resize + ToTensor on the GPU
            raw = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).to(self.device)
            return ops.pil_resize_bilinear(raw, self.size, self.size)
        img = img.resize((self.size, self.size), Image.BILINEAR)       # CPU device (host tests): torchvision Resize on PIL
        arr = np.asarray(img, dtype=np.uint8)                          # HWC
        return torch.from_numpy(arr.copy()).permute(2, 0, 1).float().div_(255.0)   # ToTensor()

    def __call__(self, prompt, n=None):
        """-> float32 [len(files) or n, 3, size, size] in [0, 1] on the device (train_pick:797-799).  With `n`, the
        file list is cycled / truncated to n images (the reference assumes len(files) == mini_num_image_per_prompt)."""
        if prompt not in self.index:
            raise KeyError(f"no reference images for prompt {prompt!r} in the index (train_pick:773,787-789)")
        if self._cache is not None and prompt in self._cache:
            t = self._cache[prompt]
        else:
            files = self.index[prompt]
            t = torch.stack([self._load(os.path.join(self.root, f)) for f in files]).to(self.device, torch.float32)
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
