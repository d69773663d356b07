"""Oracle: Flow-CPS SDE step with log-prob (`adv_grpo/diffusers_patch/
sd3_sde_with_logprob.py:77-139`, the function both training scripts import as
`sde_step_with_logprob`), plus the bf16 classifier-free-guidance combine that
precedes it (`fast.py:640-642`, `train_sd3_fast_pickscore.py:242-247`).

torch fp32 on CPU; op order follows the reference line by line.
Test infrastructure only (see oracle/__init__.py).
"""
import math
import torch


def cfg_combine(noise_pred_uncond, noise_pred_text, guidance_scale):
    """Classifier-free guidance in the dtype of the model output (fast.py:640-642; train_sd3_fast_pickscore.py:242-247)."""
    # fast.py:641-642 -- evaluated in the dtype of the transformer output (bf16 in
    # the reference run: every elementwise op rounds to bf16).
    return noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)


def sde_step_with_logprob_new(sigmas, step_index, model_output, sample, noise_level=0.7,
                              prev_sample=None, noise=None, generator=None):
    """sde.py:100-139.  `sigmas` is the scheduler's fp32 sigma table, `step_index`
    a list[int] (one per batch element, or a single entry that broadcasts, as in
    the rollout where `timestep = t.unsqueeze(0)`, fast.py:649).
    Returns (prev_sample, log_prob, prev_sample_mean, std_dev_t)."""
    model_output = model_output.float()                     # sde.py:100
    sample = sample.float()                                 # sde.py:101
    if prev_sample is not None:
        prev_sample = prev_sample.float()                   # sde.py:103
    step_index = list(step_index)
    prev_step_index = [s + 1 for s in step_index]           # sde.py:107
    shape = (-1,) + (1,) * (sample.dim() - 1)
    sigma = sigmas[step_index].view(*shape)                 # sde.py:108
    sigma_prev = sigmas[prev_step_index].view(*shape)       # sde.py:109
    std_dev_t = sigma_prev * math.sin(noise_level * math.pi / 2)        # sde.py:119
    pred_original_sample = sample - sigma * model_output                # sde.py:120
    noise_estimate = sample + model_output * (1 - sigma)                # sde.py:121
    prev_sample_mean = pred_original_sample * (1 - sigma_prev) + \
        noise_estimate * torch.sqrt(sigma_prev ** 2 - std_dev_t ** 2)   # sde.py:122
    if prev_sample is None:                                             # sde.py:125-131
        if noise is None:
            noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
        prev_sample = prev_sample_mean + std_dev_t * noise
    log_prob = -((prev_sample.detach() - prev_sample_mean) ** 2)        # sde.py:134
    log_prob = log_prob.mean(dim=tuple(range(1, log_prob.ndim)))        # sde.py:137
    return prev_sample, log_prob, prev_sample_mean, std_dev_t


def sde_step_with_logprob(sigmas, step_index, model_output, sample, noise_level=0.7, prev_sample=None, noise=None,
                          generator=None):
    """The Flow-SDE variant, sde.py:13-73 (defined next to `_new`; both training scripts import `_new`, this one is
    kept for the configs that select the true-Gaussian log-density).  Same calling convention as above.
    Returns (prev_sample, log_prob, prev_sample_mean, std_dev_t)."""
    model_output = model_output.float()                                             # sde.py:37
    sample = sample.float()                                                         # sde.py:38
    if prev_sample is not None:
        prev_sample = prev_sample.float()                                           # sde.py:40
    step_index = list(step_index)
    prev_step_index = [s + 1 for s in step_index]                                   # sde.py:43
    shape = (-1,) + (1,) * (sample.dim() - 1)
    sigma = sigmas[step_index].view(*shape)                                         # sde.py:44
    sigma_prev = sigmas[prev_step_index].view(*shape)                               # sde.py:45
    sigma_max = sigmas[1].item()                                                    # sde.py:46
    dt = sigma_prev - sigma                                                         # sde.py:47
    std_dev_t = torch.sqrt(sigma / (1 - torch.where(sigma == 1, sigma_max, sigma))) * noise_level      # sde.py:49
    prev_sample_mean = sample * (1 + std_dev_t ** 2 / (2 * sigma) * dt) + \
        model_output * (1 + std_dev_t ** 2 * (1 - sigma) / (2 * sigma)) * dt        # sde.py:53
    if prev_sample is None:                                                         # sde.py:55-62
        if noise is None:
            noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
        prev_sample = prev_sample_mean + std_dev_t * torch.sqrt(-1 * dt) * noise
    log_prob = (
        -((prev_sample.detach() - prev_sample_mean) ** 2) / (2 * ((std_dev_t * torch.sqrt(-1 * dt)) ** 2))
        - torch.log(std_dev_t * torch.sqrt(-1 * dt))
        - torch.log(torch.sqrt(2 * torch.as_tensor(math.pi)))
    )                                                                               # sde.py:64-68
    log_prob = log_prob.mean(dim=tuple(range(1, log_prob.ndim)))                    # sde.py:71
    return prev_sample, log_prob, prev_sample_mean, std_dev_t
