"""Host-side OCR reward plugin, the drop-in for `adv_grpo/ocr.py:8-65`.

The reference's `OcrScorer` is CPU work by design: PaddleOCR text recognition on uint8 NHWC images followed by a
Levenshtein distance to the quoted target text of the prompt.  Nothing in it is a GPU kernel, so it stays a host
plugin here too; this module only removes the two hard imports that make the reference file unusable without its
exact environment:
  * the recogniser is injectable (`OcrScorer(recognizer=fn)`, `fn(uint8 HxWx3 ndarray) -> [(text, confidence), ...]`);
    when none is given PaddleOCR is constructed exactly as the reference does (`ocr.py:14-19`) and a missing
    package raises at construction, never silently scores zero;
  * the edit distance is computed here (`levenshtein`, unit insert / delete / substitute costs = `Levenshtein.distance`).
Reward arithmetic, text normalisation, the containment shortcut, the per-image failure penalty and the return type
(`list[float]`) follow `ocr.py:31-65` line by line.
"""
import numpy as np


def levenshtein(a, b):
    """Classic edit distance with unit costs (python-Levenshtein `distance`, used at ocr.py:50)."""
    if a == b:
        return 0
    if len(a) < len(b):
        a, b = b, a
    if not b:
        return len(a)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def _paddle_recognizer(use_gpu):
    try:
        from paddleocr import PaddleOCR
    except Exception as e:
        raise ImportError("OcrScorer needs a text recogniser: install paddleocr (as the reference does, "
                          f"adv_grpo/ocr.py:1) or pass recognizer=callable ({e})")
    ocr = PaddleOCR(use_angle_cls=False, lang="en", use_gpu=use_gpu, show_log=False)

    def recognize(img):
        result = ocr.ocr(img, cls=False)
        return [(res[1][0], res[1][1]) for res in result[0]] if result and result[0] else []

    return recognize


class OcrScorer:
    def __init__(self, use_gpu=False, recognizer=None):
        self.recognize = recognizer if recognizer is not None else _paddle_recognizer(use_gpu)

    def __call__(self, images, prompts):
        prompts = [prompt.split('"')[1] for prompt in prompts]              # the quoted target text (ocr.py:31)
        assert len(images) == len(prompts), "Images and prompts must have the same length"
        rewards = []
        for img, prompt in zip(images, prompts):
            if not isinstance(img, np.ndarray):                              # PIL image
                img = np.array(img)
            try:
                lines = self.recognize(img)
                text = "".join(t if conf > 0 else "" for t, conf in lines)
                text = text.replace(" ", "").lower()
                prompt = prompt.replace(" ", "").lower()
                dist = 0 if prompt in text else levenshtein(text, prompt)
                if dist > len(prompt):                                       # many unrelated characters: cap the penalty
                    dist = len(prompt)
            except Exception as e:                                           # recogniser failure = maximum penalty
                print(f"OCR processing failed: {str(e)}")
                dist = len(prompt)
            rewards.append(1 - dist / len(prompt))
        return rewards
