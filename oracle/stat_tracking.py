"""Oracle: group-relative advantage (`adv_grpo/stat_tracking.py:18-47`,
type='grpo') and the zero-std statistics (`train_sd3_fast_pickscore.py:195-229`).
numpy float64, following the reference's order of operations.
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np


def grpo_advantages(prompts, rewards, global_std):
    """stat_tracking.py:18-47 with an empty history (the scripts clear the tracker
    every epoch, train_sd3_fast_pickscore.py:989)."""
    prompts = np.array(prompts)
    rewards = np.array(rewards, dtype=np.float64)
    adv = np.zeros_like(rewards)
    for p in np.unique(prompts):
        sel = prompts == p
        grp = rewards[sel]
        mean = np.mean(grp, axis=0, keepdims=True)
        if global_std:
            std = np.std(rewards, axis=0, keepdims=True) + 1e-4
        else:
            std = np.std(grp, axis=0, keepdims=True) + 1e-4
        adv[sel] = (grp - mean) / std
    return adv


def zero_std_ratio(prompts, ori_avg):
    """train_sd3_fast_pickscore.py:195-229."""
    prompts = np.array(prompts)
    ori_avg = np.asarray(ori_avg)
    stds = np.array([np.std(ori_avg[prompts == p]) for p in np.unique(prompts)])
    return np.count_nonzero(stds == 0) / len(stds), stds.mean()


def mode_advantages(prompts, rewards, mode):
    """The other `type`s of PerPromptStatTracker.update, stat_tracking.py:48-70, empty history:
    'rwr' returns the rewards (:48-50); 'sft' marks the elements equal to the maximum over the group's whole
    block (:52-53, torch.max over all elements); 'dpo' puts +1 on the group's first arg-max and -1 on its first
    arg-min, and on members 1 / 0 when the whole group is tied (:54-68; 1-D rewards)."""
    prompts = np.array(prompts)
    rewards = np.array(rewards, dtype=np.float64)
    adv = np.zeros_like(rewards)
    for p in np.unique(prompts):
        sel = prompts == p
        grp = rewards[sel]
        if mode == "rwr":
            adv[sel] = grp
        elif mode == "sft":
            adv[sel] = (grp == grp.max()).astype(np.float64)
        elif mode == "dpo":
            hi, lo = int(np.argmax(grp)), int(np.argmin(grp))       # first occurrence, like torch.argmax / argmin
            if hi == lo:
                lo, hi = 0, 1
            out = np.zeros(len(grp))
            out[hi], out[lo] = 1.0, -1.0
            adv[sel] = out
        else:
            raise ValueError(mode)
    return adv
