"""Drop-in for `adv_grpo/diffusers_patch/sd3_sde_with_logprob.py` (the reference file both training
scripts import `sde_step_with_logprob_new` from, train_sd3_fast_pickscore.py:21, fast.py:9).
Same name, arguments and return convention; the arithmetic is one fused sm_100a kernel
(`advgrpo_cfg_sde_step_logprob`) with the scheduler lookup done on the device.
"""
import itertools
import threading

import torch

from .. import ops

_offset_lock = threading.Lock()
_offset = itertools.count()


def _next_offset(n_elems):
    # each call consumes a disjoint slice of the Philox counter space
    with _offset_lock:
        k = next(_offset)
    return k * (1 << 34)


def _as_bf16_pair(model_output, sample):
    return model_output.to(torch.bfloat16), sample.to(torch.bfloat16)


def sde_step_with_logprob_new(self, model_output, timestep, sample, noise_level=0.7, prev_sample=None,
                              generator=None, noise=None, _variant=ops.SDE_FLOW_CPS):
    """Flow-CPS reverse step.  Returns (prev_sample, log_prob, prev_sample_mean, std_dev_t) like
    sde.py:139.  `log_prob` is differentiable w.r.t. `model_output` when `prev_sample` is given
    (the replay of train_sd3_fast_pickscore.py:258-267).  Extra keyword `noise` injects the Gaussian
    draw (parity runs); otherwise it is generated in-kernel (Philox4x32-10) keyed by `generator`'s
    seed, or drawn with torch.randn(generator=...) when `generator` is a CUDA generator whose exact
    stream is wanted (`generator.advgrpo_exact = True`)."""
    dev = sample.device
    mo, x = _as_bf16_pair(model_output, sample)
    shape = (-1,) + (1,) * (sample.dim() - 1)
    if torch.is_tensor(timestep):
        t = timestep.reshape(-1).to(torch.float32)
    else:
        t = torch.tensor([float(timestep)], dtype=torch.float32)
    sched_t = self.timesteps
    sigmas = self.sigmas
    if prev_sample is not None:
        logp, mean, std = ops.sde_logprob_replay(mo, x, prev_sample.to(torch.bfloat16), t, sched_t, sigmas, 1.0,
                                                 noise_level, cfg=False, want_mean=True, variant=_variant)
        return prev_sample.float(), logp, mean, std.view(*shape)
    seed = 0
    if generator is not None:
        if getattr(generator, "advgrpo_exact", False) and noise is None:
            noise = torch.randn(model_output.shape, generator=generator, device=dev, dtype=torch.float32)
        seed = generator.initial_seed()
    else:
        seed = torch.initial_seed()
    prev, logp, mean, std = ops.cfg_sde_step_logprob(None, mo, x, t, sched_t, sigmas, 1.0, noise_level,
                                                     noise=noise, seed=seed, offset=_next_offset(x.numel()),
                                                     want_mean=True, variant=_variant)
    return prev.float(), logp, mean, std.view(*shape)


def sde_step_with_logprob(self, model_output, timestep, sample, noise_level=0.7, prev_sample=None, generator=None,
                          noise=None):
    """Flow-SDE reverse step with the Gaussian log-density (sde.py:13-73): same signature and return convention; the
    same fused kernel in its second variant.  (Both training scripts import `sde_step_with_logprob_new` under this
    name, train_sd3_fast_pickscore.py:21; a module-level import of `sde_step_with_logprob` gets this function, as
    in the reference file.)"""
    return sde_step_with_logprob_new(self, model_output, timestep, sample, noise_level, prev_sample, generator, noise,
                                     _variant=ops.SDE_FLOW_SDE)
