"""Oracle: GRPO clipped policy-gradient loss and its logged statistics
(`scripts/train_sd3_fast_pickscore.py:1111-1162`).  The advantages arrive as
float64 (quirk Q5), so the loss is float64.
Test infrastructure only (see oracle/__init__.py).
"""
import torch


def grpo_clip_loss(log_prob, old_log_prob, advantages, clip_range, adv_clip_max):
    """Clipped GRPO loss and its logged statistics (train_sd3_fast_pickscore.py:1111-1162)."""
    adv = torch.clamp(advantages, -adv_clip_max, adv_clip_max)          # :1111-1115
    ratio = torch.exp(log_prob - old_log_prob)                          # :1116
    unclipped = -adv * ratio                                            # :1117
    clipped = -adv * torch.clamp(ratio, 1.0 - clip_range, 1.0 + clip_range)   # :1118-1122
    policy_loss = torch.mean(torch.maximum(unclipped, clipped))         # :1123
    info = {
        "approx_kl": 0.5 * torch.mean((log_prob - old_log_prob) ** 2),               # :1132-1135
        "clipfrac": torch.mean((torch.abs(ratio - 1.0) > clip_range).float()),       # :1136-1142
        "clipfrac_gt_one": torch.mean((ratio - 1.0 > clip_range).float()),           # :1143-1149
        "clipfrac_lt_one": torch.mean((1.0 - ratio > clip_range).float()),           # :1150-1156
        "policy_loss": policy_loss,
    }
    return policy_loss, info


def kl_loss(prev_sample_mean, prev_sample_mean_ref):
    """KL regulariser of the beta > 0 branch (`train_sd3_fast_pickscore.py:1124-1128`; the 1 / (2 std^2) factor is
    commented out in the reference): mean over the batch of the per-sample mean squared distance between the
    prev_sample_mean of the LoRA model and of the adapter-disabled model."""
    kl = ((prev_sample_mean - prev_sample_mean_ref) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    return torch.mean(kl)
