"""Extracts the call signatures of the reference's boundary symbols by parsing the
reference sources with `ast` (nothing is imported or executed: most reference modules cannot be imported here)
and writes them to tests/golden/signatures.json.  Run in the build container, where /root/reference exists:

    python tests/golden/make_signatures.py

tests/test_host_logic.py::test_boundary_signatures_match_reference then checks that every mirror in
adv_grpo_b200 accepts the same parameters (names, order, defaults); mirrors may add trailing optional ones."""
import ast
import json
import os

REF = os.environ.get("ADVGRPO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

# (reference file, dotted path of the def inside it: Class.method / factory._fn)
SYMBOLS = [
    ("adv_grpo/stat_tracking.py", "PerPromptStatTracker.__init__"),
    ("adv_grpo/stat_tracking.py", "PerPromptStatTracker.update"),
    ("adv_grpo/stat_tracking.py", "PerPromptStatTracker.get_stats"),
    ("adv_grpo/stat_tracking.py", "PerPromptStatTracker.clear"),
    ("adv_grpo/ema.py", "EMAModuleWrapper.__init__"),
    ("adv_grpo/ema.py", "EMAModuleWrapper.step"),
    ("adv_grpo/ema.py", "EMAModuleWrapper.copy_ema_to"),
    ("adv_grpo/ema.py", "EMAModuleWrapper.copy_temp_to"),
    ("adv_grpo/ema.py", "EMAModuleWrapper.state_dict"),
    ("adv_grpo/ema.py", "EMAModuleWrapper.load_state_dict"),
    ("adv_grpo/pickscore_scorer.py", "PickScoreScorer.__init__"),
    ("adv_grpo/pickscore_scorer.py", "PickScoreScorer.__call__"),
    ("adv_grpo/pick_score_training.py", "CLIPCriterion.__init__"),
    ("adv_grpo/pick_score_training.py", "CLIPCriterion.forward"),
    ("adv_grpo/pick_score_training.py", "CLIPCriterion.calc_loss"),
    ("adv_grpo/diffusers_patch/sd3_sde_with_logprob.py", "sde_step_with_logprob"),
    ("adv_grpo/diffusers_patch/sd3_sde_with_logprob.py", "sde_step_with_logprob_new"),
    ("adv_grpo/diffusers_patch/sd3_pipeline_with_logprob_fast.py", "pipeline_with_logprob_random"),
    ("adv_grpo/diffusers_patch/train_dreambooth_lora_sd3.py", "encode_prompt"),
    ("adv_grpo/ocr.py", "OcrScorer.__init__"),
    ("adv_grpo/ocr.py", "OcrScorer.__call__"),
    ("adv_grpo/rewards.py", "multi_score"),
    ("adv_grpo/rewards.py", "multi_score._fn"),
    ("adv_grpo/rewards.py", "pickscore_score"),
    ("adv_grpo/rewards.py", "pickscore_score._fn"),
    ("adv_grpo/rewards.py", "pickscore_cotrain_score"),
    ("adv_grpo/rewards.py", "pickscore_cotrain_score._fn"),
    ("adv_grpo/rewards.py", "dino_patch_cotrain_score"),
    ("adv_grpo/rewards.py", "dino_patch_cotrain_score._fn"),
    ("adv_grpo/rewards.py", "ocr_score"),
    ("adv_grpo/rewards.py", "ocr_score._fn"),
]


def find(tree, dotted):
    node = tree
    for part in dotted.split("."):
        for child in ast.walk(node) if node is not tree else node.body:
            if isinstance(child, (ast.FunctionDef, ast.ClassDef)) and child.name == part and child is not node:
                node = child
                break
        else:
            raise KeyError(dotted)
    return node


def signature(fn):
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    defaults = [ast.unparse(d) for d in a.defaults]
    n_req = len(pos) - len(defaults)
    params = [{"name": n, "kind": "positional", "default": None if i < n_req else defaults[i - n_req]}
              for i, n in enumerate(pos)]
    for x, d in zip(a.kwonlyargs, a.kw_defaults):
        params.append({"name": x.arg, "kind": "keyword_only", "default": None if d is None else ast.unparse(d)})
    return {"params": params, "varargs": a.vararg is not None, "varkw": a.kwarg is not None, "line": fn.lineno}


def main():
    out = {}
    for rel, dotted in SYMBOLS:
        with open(os.path.join(REF, rel)) as f:
            tree = ast.parse(f.read())
        out[f"{rel}::{dotted}"] = signature(find(tree, dotted))
    # the registry keys of multi_score (rewards.py:1014-1038)
    with open(os.path.join(REF, "adv_grpo/rewards.py")) as f:
        tree = ast.parse(f.read())
    ms = find(tree, "multi_score")
    keys = []
    for node in ast.walk(ms):
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "score_functions" for t in node.targets):
            keys = [ast.literal_eval(k) for k in node.value.keys]
    out["adv_grpo/rewards.py::multi_score.score_functions"] = {"keys": keys}
    with open(os.path.join(HERE, "signatures.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {len(out)} entries")


if __name__ == "__main__":
    main()
