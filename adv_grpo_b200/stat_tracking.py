"""Drop-in for `adv_grpo/stat_tracking.py` (`PerPromptStatTracker`).  Same constructor,
`update(prompts, rewards) -> float64 ndarray`, `get_stats()`, `clear()`; the arithmetic is the
group-advantage kernel.  `update_device` is the B200-native entry: it takes the gathered reward
tensor and the gathered prompt-id rows ON THE DEVICE and returns device tensors, skipping the
tokenizer decode and both host round trips of `train_sd3_fast_pickscore.py:931,962-970,995-999`.
History semantics: the scripts clear the tracker every epoch (`:989`), so an update only ever sees
the current epoch's rewards; carrying statistics across un-cleared updates is not supported
(raises).  The 'rwr' / 'sft' / 'dpo' modes (unused by the scripts) are modes of the same kernel.  Rewards enter the kernel
as float32 (the scripts' reward tensors are float32; a float64 ndarray passed to `update` is cast), the statistics and the
advantages are float64 like the reference's numpy arithmetic."""
import zlib

import numpy as np
import torch

from . import ops


class PerPromptStatTracker:
    def __init__(self, global_std=False, device="cuda"):
        self.global_std = global_std
        self.device = device
        self.stats = {}
        self.history_prompts = set()
        self._last = None
        self._pending_keys = None

    def update_device(self, prompt_keys, rewards, type="grpo"):
        """prompt_keys int64 [N] or [N, L] (e.g. gathered `prompt_ids`), rewards f32 [N] or [N, T], both CUDA.
        Returns advantages f64 (same shape as rewards) on the device."""
        adv, stats = ops.group_advantage(rewards, prompt_keys, self.global_std, want_stats=True, mode=type)
        self._last = stats
        self._pending_keys = prompt_keys          # digested lazily in get_stats(): no host sync on the update path
        return adv

    def _absorb_pending_keys(self):
        """`trained_prompt_num` of the reference = number of distinct prompts ever seen (`history_prompts`, :36-37): one
        64-bit digest per key row, computed on the device and read back only when the statistics are asked for."""
        k = getattr(self, "_pending_keys", None)
        if k is None:
            return
        self._pending_keys = None
        k = k.reshape(k.shape[0], -1).to(torch.int64)
        w = torch.arange(1, k.shape[1] + 1, device=k.device, dtype=torch.int64) * 0x9E3779B1 + 0x7F4A7C15
        self.history_prompts.update((k * w).sum(dim=1).unique().tolist())

    def update(self, prompts, rewards, type="grpo"):
        if type not in ops.ADV_MODES:
            raise ValueError(f"unknown type {type!r}; the reference knows 'grpo', 'rwr', 'sft', 'dpo'")
        if self.stats:
            raise NotImplementedError("call clear() between updates (the scripts do, train_sd3_fast_pickscore.py:989)")
        prompts = list(prompts)
        keys = torch.tensor([[zlib.crc32(p.encode()), zlib.adler32(p.encode()), len(p)] for p in prompts],
                            dtype=torch.int64, device=self.device)
        r = torch.as_tensor(np.asarray(rewards, dtype=np.float64), dtype=torch.float32, device=self.device)
        adv = self.update_device(keys, r, type)
        self._pending_keys = None                 # host path: the prompt strings themselves are the history keys
        uniq = set(prompts)
        for p in uniq:
            self.stats[p] = prompts.count(p)
            self.history_prompts.add(hash(p))
        return adv.cpu().numpy()

    def get_stats(self):
        self._absorb_pending_keys()
        if self._last is not None and not self.stats:
            s = self._last.tolist()
            return s[1], len(self.history_prompts)
        avg = sum(self.stats.values()) / len(self.stats) if self.stats else 0
        return avg, len(self.history_prompts)

    def zero_std_stats(self):
        """(zero_std_ratio, reward_std_mean) of the last update (calculate_zero_std_ratio, :195-229)."""
        s = self._last.tolist()
        return s[2], s[3]

    def clear(self):
        self.stats = {}
