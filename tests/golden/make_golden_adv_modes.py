"""Golden vectors G12 for the non-'grpo' modes of PerPromptStatTracker.update (`type` in {'rwr', 'sft', 'dpo'},
adv_grpo/stat_tracking.py:48-70), produced by EXECUTING the reference file verbatim.  Build container only:

    python tests/golden/make_golden_adv_modes.py        # writes tests/golden/golden_adv_modes.json
"""
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_stat_tracking", f"{REF}/adv_grpo/stat_tracking.py")
    st = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(st)
    rng = np.random.RandomState(12)
    n_prompts, G = 5, 6
    prompts = [f"prompt {i}" for i in range(n_prompts) for _ in range(G)]
    perm = rng.permutation(len(prompts))
    prompts = [prompts[i] for i in perm]
    rewards = np.round(rng.rand(len(prompts)), 1).astype(np.float32)          # one decimal: ties inside groups
    rewards[[i for i, p in enumerate(prompts) if p == "prompt 2"]] = 0.5      # an all-equal group
    gold = {"prompts": prompts, "rewards": rewards.tolist()}
    for mode in ("rwr", "sft", "dpo"):
        for gs in (False, True):
            tr = st.PerPromptStatTracker(global_std=gs)
            gold[f"{mode}_global{int(gs)}"] = tr.update(prompts, rewards, type=mode).tolist()
    # 2-D rewards [N, T] (the scripts' layout): 'sft' compares with the maximum over the WHOLE [n, T] block of a group
    rew2 = np.stack([rewards, np.round(rng.rand(len(prompts)), 1).astype(np.float32)], axis=1)
    gold["rewards_2d"] = rew2.tolist()
    for mode in ("rwr", "sft"):
        tr = st.PerPromptStatTracker(global_std=True)
        gold[f"{mode}_2d"] = tr.update(prompts, rew2, type=mode).tolist()
    with open(os.path.join(OUT, "golden_adv_modes.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print({k: (len(v) if isinstance(v, list) else v) for k, v in gold.items()})


if __name__ == "__main__":
    main()
