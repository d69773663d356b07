"""Drop-in for the plugin registry of `adv_grpo/rewards.py`: `multi_score(device, score_dict)` returns
`_fn(images, prompts, metadata, scorer=None, ref_images=None, only_strict=True, head=None, fusion=None,
layer_ids=None, temperature=0.2) -> (score_details, {})` with `score_details[name]` per reward and
the weighted sum in `score_details['avg']` (`rewards.py:1012-1095`).  Factory protocol unchanged:
`factory(device)` if it takes `device`, else `factory()`; each returns `_fn(...) -> (scores, info)`.

On the hot path (B200 kernels, device-resident, no host round trip):
  pickscore_cotrain, pickscore, dino_patch_cotrain
Host plugin kept as in the reference:  ocr (`ocr.OcrScorer`: PaddleOCR or an injected recogniser on the CPU +
Levenshtein reward; a missing recogniser raises at construction).
All other registry keys of the reference exist so configs do not KeyError; they raise
NotImplementedError when instantiated (remote HTTP / sglang / SigLIP / aesthetic ... scorers are
outside the hot path).
Scores stay CUDA tensors (the scripts consume them with `torch.as_tensor(value).float()`,
train_sd3_fast_pickscore.py:849-856); the weighted sum is a device axpy, not a Python list loop.
"""
import torch

from . import ops


def _to_unit_tensor(images, device):
    if not torch.is_tensor(images):
        images = torch.as_tensor(images)
    images = images.to(device)
    if images.dtype == torch.uint8:
        return images if images.shape[1] == 3 else images.permute(0, 3, 1, 2).contiguous()
    return images


PICKSCORE_KWARGS = {}      # extra PickScoreScorer kwargs of the frozen `pickscore` reward (state_dict=..., cfg=...)


def pickscore_score(device):
    from .pickscore_scorer import PickScoreScorer
    scorer = PickScoreScorer(dtype=torch.float32, device=device, **PICKSCORE_KWARGS)   # own frozen model (rewards.py:564)

    def _fn(images, prompts, metadata):
        return scorer(prompts, _to_unit_tensor(images, device)), {}

    return _fn


def pickscore_cotrain_score(device):
    def _fn(scorer, images, prompts, metadata):
        return scorer(prompts, _to_unit_tensor(images, device)), {}

    return _fn


def dino_patch_cotrain_score(device, n_patches=64):
    from .dinov2 import dino_patch_scores

    def _fn(scorer, head, images, prompts, metadata, cls_weight=0.7):
        images = _to_unit_tensor(images, device)
        if images.dtype == torch.uint8:
            images = images.float() / 255.0
        pix = ops.dino_preprocess(images, 518)                        # rewards.py:379-391 fused
        with torch.no_grad():
            feats = scorer.forward_features(pix)                       # [B, N+1, D]
        N = feats.shape[1] - 1
        n_select = min(n_patches, N)
        idx = torch.randint(0, N, (feats.shape[0], n_select), device=feats.device)   # rewards.py:406
        # gather + L2-norm (+1e-6) + DINOHead + 0.7 / 0.3 mix on the native kernels (rewards.py:408-419)
        hybrid, cls_score, patch_scores = dino_patch_scores(head, feats, idx, cls_weight)
        return hybrid, {"cls_score": cls_score, "patch_scores": patch_scores, "patch_indices": idx,
                        "cls_weight": cls_weight}

    return _fn


OCR_KWARGS = {}            # extra OcrScorer kwargs (recognizer=callable replaces PaddleOCR)


def ocr_score(device):
    from .ocr import OcrScorer                                 # host plugin: PaddleOCR (or OCR_KWARGS["recognizer"]) on CPU
    scorer = OcrScorer(**OCR_KWARGS)

    def _fn(images, prompts, metadata):
        if torch.is_tensor(images):
            images = (images * 255).round().clamp(0, 255).to(torch.uint8).cpu().numpy().transpose(0, 2, 3, 1)
        return scorer(images, prompts), {}

    return _fn


def _out_of_scope(name):
    def factory(device=None):
        raise NotImplementedError(f"reward '{name}' is outside the B200 hot path (SURVEY.md section 2.1 #4); use the "
                                  "reference implementation for it")
    factory.__name__ = name
    return factory


score_functions = {
    "deqa": _out_of_scope("deqa"), "ocr": ocr_score, "video_ocr": _out_of_scope("video_ocr"),
    "imagereward": _out_of_scope("imagereward"), "pickscore": pickscore_score, "qwenvl": _out_of_scope("qwenvl"),
    "aesthetic": _out_of_scope("aesthetic"), "jpeg_compressibility": _out_of_scope("jpeg_compressibility"),
    "unifiedreward": _out_of_scope("unifiedreward"), "geneval": _out_of_scope("geneval"),
    "clipscore": _out_of_scope("clipscore"), "image_similarity": _out_of_scope("image_similarity"),
    "image_similarity_eval": _out_of_scope("image_similarity_eval"),
    "constractive_external": _out_of_scope("constractive_external"), "discriminator": _out_of_scope("discriminator"),
    "pickscore_cotrain": pickscore_cotrain_score, "pickscore_patch": _out_of_scope("pickscore_patch"),
    "dino_cotrain": _out_of_scope("dino_cotrain"), "dino_multi_cotrain": _out_of_scope("dino_multi_cotrain"),
    "dino_patch_cotrain": dino_patch_cotrain_score, "siglip_cotrain": _out_of_scope("siglip_cotrain"),
    "siglip_image_similarity": _out_of_scope("siglip_image_similarity"),
}


def multi_score(device, score_dict):
    score_fns = {}
    for name, weight in score_dict.items():
        factory = score_functions[name]
        score_fns[name] = factory(device) if "device" in factory.__code__.co_varnames else factory()

    def _fn(images, prompts, metadata, scorer=None, ref_images=None, only_strict=True, head=None, fusion=None,
            layer_ids=None, temperature=0.2):
        total = None
        score_details = {}
        for name, weight in score_dict.items():
            if name == "pickscore_cotrain":
                scores, info = score_fns[name](scorer, images, prompts, metadata)
            elif name in ("dino_patch_cotrain", "dino_cotrain", "siglip_cotrain"):
                scores, info = score_fns[name](scorer, head, images, prompts, metadata)
            else:
                scores, info = score_fns[name](images, prompts, metadata)
            score_details[name] = scores
            s = torch.as_tensor(scores, device=device).float() if not torch.is_tensor(scores) else scores.float()
            total = weight * s if total is None else total + weight * s
        score_details["avg"] = total
        return score_details, {}

    return _fn
