"""Drop-in for `adv_grpo/pickscore_scorer.py`: `PickScoreScorer(device, dtype)` exposing `.processor`
(with `.tokenizer`), `.model` and `__call__(prompt, images) -> scores` (cosine * exp(logit_scale) / 26).

B200-native path: images that are already CUDA tensors in [0,1] (what the rollout produces) are
quantised, PIL-exact bicubic-resized to 224 and normalised by one preprocessing kernel chain -- no
device->host->PIL->device trip (`adv_grpo/rewards.py:581-584`); the text tower runs ONCE per distinct
prompt (the reference recomputes it for each of the G identical prompts); the score is a fused
row-dot instead of the diagonal of a [B,B] matmul (`pickscore_scorer.py:47-51`).

No tokenizer files or pretrained weights exist on the box: `processor.tokenizer` is a deterministic
synthetic CLIP-shaped tokenizer (BOS, hashed word ids, EOS) unless a real one is supplied, and the
model is seeded-random CLIP-ViT-H/14 unless a state dict is supplied.
"""
import os
import threading
import zlib

import numpy as np
import torch

from . import ops
from .clip import CLIPModel
from .weights import CLIP_H, init_clip


class _Encoding(dict):
    """`tokenizer(...)` result with both item and attribute access (`.input_ids`), like transformers' BatchEncoding."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class SyntheticCLIPTokenizer:
    bos, eos, vocab = 49406, 49407, 49408

    def __init__(self, vocab=49408):
        self.vocab = vocab
        self.bos, self.eos = vocab - 2, vocab - 1

    def _ids(self, text, max_length):
        words = text.lower().split()
        ids = [self.bos] + [1 + zlib.crc32(w.encode()) % (self.vocab - 3) for w in words][: max_length - 2] + [self.eos]
        return ids

    def __call__(self, text, padding=True, truncation=True, max_length=77, return_tensors="pt", **_):
        if isinstance(text, str):
            text = [text]
        rows = [self._ids(t, max_length) for t in text]
        width = max_length if padding == "max_length" else max(len(r) for r in rows)
        out = np.zeros((len(rows), width), dtype=np.int64)
        for i, r in enumerate(rows):
            out[i, :len(r)] = r
        return _Encoding({"input_ids": torch.from_numpy(out), "attention_mask": torch.from_numpy((out != 0).astype(np.int64))})

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(str(int(t)) for t in row if int(t) not in (0, self.bos, self.eos)) for row in ids]


class CLIPProcessorLike:
    """`processor(images=...)` / `processor(text=...)` / `processor.tokenizer(...)`."""

    def __init__(self, tokenizer, device, size=224):
        self.tokenizer, self.device, self.size = tokenizer, device, size

    def __call__(self, images=None, text=None, return_tensors="pt", **kw):
        if text is not None:
            return self.tokenizer(text, **kw)
        return {"pixel_values": images_to_pixel_values(images, self.device, self.size)}


def images_to_pixel_values(images, device, size=224, dtype=torch.bfloat16):
    """CUDA float tensor in [0,1] (bf16 quantisation semantics of rewards.py:581), uint8 tensor, or a list
    of PIL images / HWC uint8 arrays (the reference's inputs) -> CLIP pixel_values on the device."""
    if torch.is_tensor(images):
        t = images.to(device)
        t = t if t.dtype == torch.uint8 else t.to(torch.bfloat16)
    else:
        arr = np.stack([np.asarray(im) for im in images])            # [B,H,W,3] uint8
        t = torch.from_numpy(arr).to(device).permute(0, 3, 1, 2).contiguous()
    return ops.clip_preprocess(t, size, dtype=dtype)


# Reward worker threads (train_sd3_fast_pickscore.py:668) may capture the image-tower graph; set to False to let only
# the main thread capture (workers then run the tower eagerly until the main thread has captured it).
CAPTURE_FROM_ANY_THREAD = os.environ.get("ADVGRPO_SCORER_CAPTURE_MAIN_ONLY", "0") != "1"


class _GraphedImageTower:
    """CUDA-graph replay of `model.get_image_features` for a fixed batch shape (the ViT-H forward is ~400
    launches of a few microseconds each: launch-bound when issued eagerly).  Re-captured whenever a
    parameter version changes (discriminator steps update the last vision blocks in place)."""

    def __init__(self, model):
        self.model, self.entries = model, {}

    def _version(self):
        return sum(p._version for p in self.model.vision_model.parameters()) + self.model.visual_projection.weight._version

    def __call__(self, pixel_values):
        from . import _lib
        key = (tuple(pixel_values.shape), pixel_values.dtype)
        ver = self._version()
        ent = self.entries.get(key)
        if ent is None or ent[0] != ver:
            if ent is None or ent[1] is not None:            # first sight (or stale graph): run eagerly once to warm up
                self.entries[key] = (ver, None, None, None, 0)
                return self.model.get_image_features(pixel_values=pixel_values)
        if ent[1] is None:
            if not CAPTURE_FROM_ANY_THREAD and threading.current_thread() is not threading.main_thread():
                return self.model.get_image_features(pixel_values=pixel_values)
            static_in = pixel_values.clone()
            # warm up in THIS thread right before capturing: per-thread lazily created state (cuBLAS handle of the
            # projection, per-thread workspaces) must exist, creating it inside a capture invalidates the capture
            self.model.get_image_features(pixel_values=static_in)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            # thread_local: a capture started from a reward worker thread must not turn the sampling thread's
            # allocations / event queries into capture errors (the default "global" mode does)
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                out = self.model.get_image_features(pixel_values=static_in)
            ent = (ver, g, static_in, out, _lib.launch_count() - n0)
            self.entries[key] = ent
        _, g, static_in, out, n_kernels = ent
        static_in.copy_(pixel_values)
        g.replay()
        _lib.add_launches(n_kernels)
        return out.clone()


class PickScoreScorer(torch.nn.Module):
    def __init__(self, device="cuda", dtype=torch.float32, cfg=CLIP_H, state_dict=None, tokenizer=None, seed=3,
                 use_cuda_graph=True, reference_score_arithmetic=False):
        super().__init__()
        self.use_cuda_graph = use_cuda_graph
        # Quirk Q10: the reference scorer, instantiated in bf16 by the training script (train_pick:657 passes
        # inference_dtype), L2-normalises the features, takes the text-image dot product and applies logit_scale / 26
        # IN bf16 (adv_grpo/pickscore_scorer.py:40-51) and returns bf16 scores (3 significant digits).  The default here
        # does that tail in fp32 on the same bf16 features (scores differ by <= 1 bf16 ulp of the reference's);
        # reference_score_arithmetic=True reproduces the bf16 rounding sequence and dtype.
        self.reference_score_arithmetic = reference_score_arithmetic
        self.device, self.dtype = device, dtype
        if state_dict is None:
            state_dict = init_clip(cfg, seed=seed, device=device, dtype=torch.bfloat16)
        # the kernels compute in bf16 (fp32 accumulation); dtype is recorded for API parity
        self.model = CLIPModel.from_params(state_dict, cfg, device=device, dtype=torch.bfloat16)
        self.processor = CLIPProcessorLike(tokenizer or SyntheticCLIPTokenizer(cfg["vocab"]), device, cfg["image"])
        self._image_tower = None
        self._text_cache = {}
        self._lock = threading.RLock()

    @torch.no_grad()
    def __call__(self, prompt, images):
        # The training scripts call the reward function from an 8-worker thread pool while the main thread keeps
        # sampling (train_sd3_fast_pickscore.py:668,816-817).  The graphed image tower owns static buffers and the
        # text cache is shared, so calls on one scorer are serialised (their GPU work is stream-ordered anyway).
        with self._lock:
            return self._score(prompt, images)

    def _score(self, prompt, images):
        model = self.model.module if hasattr(self.model, "module") else self.model   # quirk Q3 tolerated
        pixel_values = images_to_pixel_values(images, self.device, self.processor.size)
        if isinstance(prompt, str):
            prompt = [prompt]
        uniq = list(dict.fromkeys(prompt))
        if self.use_cuda_graph and torch.device(self.device).type == "cuda":
            if self._image_tower is None or self._image_tower.model is not model:
                self._image_tower = _GraphedImageTower(model)
            image_embs = self._image_tower(pixel_values)
        else:
            image_embs = model.get_image_features(pixel_values=pixel_values)
        image_embs = image_embs.to(torch.bfloat16)
        self._last_image_feats_bf16 = image_embs
        # the text tower is frozen: one forward per distinct prompt, cached across calls (generated and
        # reference images of a group share the prompt; the reference recomputes it for every image)
        tver = sum(p._version for p in model.text_model.parameters())
        feats = []
        for pr in uniq:
            hit = self._text_cache.get(pr)
            if hit is None or hit[0] != tver:
                ids = self.processor.tokenizer([pr], padding=True, truncation=True, max_length=77)["input_ids"].to(self.device)
                if len(self._text_cache) > 4096:
                    self._text_cache.clear()
                hit = (tver, model.get_text_features(input_ids=ids).to(torch.bfloat16))
                self._text_cache[pr] = hit
            feats.append(hit[1])
        text_embs = feats[0] if len(feats) == 1 else torch.cat(feats, 0)
        index = None if len(uniq) == 1 else torch.tensor([uniq.index(p) for p in prompt], device=self.device)
        # score tail (pickscore_scorer.py:44-51) as ONE kernel: L2 norms, row dot, exp(logit_scale) / 26 -- in fp32 on the
        # bf16 tower outputs by default, or with the reference's bf16 rounding sequence and dtype (quirk Q10)
        scores = ops.pickscore_head(image_embs, text_embs, index, model.logit_scale, self.reference_score_arithmetic)
        return scores.to(torch.bfloat16) if self.reference_score_arithmetic else scores
