"""Oracle: the rollout loop (`adv_grpo/diffusers_patch/sd3_pipeline_with_logprob_fast.py:551-674`,
`pipeline_with_logprob_random`) and the replay micro-step (`scripts/train_sd3_fast_pickscore.py:233-267,
1111-1123`) on top of the other oracle pieces.  `bf16_io=True` reproduces the reference's dtype
boundaries (bf16 transformer I/O and CFG arithmetic, bf16 stored latents, fp32 SDE math); with
`bf16_io=False` everything is fp32 (BASELINE config 1, "CPU fp32 plumbing").
Test infrastructure only (see oracle/__init__.py)."""
import torch

from . import grpo_loss, sde, vae
from .scheduler import FlowMatchEulerOracle


def rollout(mmdit, vae_params, prompt_embeds, pooled, neg_embeds, neg_pooled, latents, num_steps, guidance_scale,
            noise_level, train_num_steps, random_timestep=0, noises=None, bf16_io=True, decode=True):
    """prompt_embeds [G, n_txt, J] etc. already repeated G times; latents [G, 16, h, w]; noises: list of
    per-step fp32 noise tensors (injected draw) or None."""
    sch = FlowMatchEulerOracle()
    timesteps = sch.set_timesteps(num_steps)
    cast = (lambda t: t.bfloat16()) if bf16_io else (lambda t: t.float())
    latents = cast(latents)
    embeds = torch.cat([neg_embeds, prompt_embeds], 0)
    pooled_all = torch.cat([neg_pooled, pooled], 0)
    all_latents, all_log_probs, all_timesteps = [], [], []
    for i, t in enumerate(timesteps):
        in_window = random_timestep <= i < random_timestep + train_num_steps
        cur = noise_level if in_window else 0
        if i == random_timestep:
            all_latents.append(latents)
        x_in = torch.cat([latents] * 2)
        pred = mmdit.forward(x_in.float(), t.expand(x_in.shape[0]), embeds.float(), pooled_all.float())
        pred = cast(pred)
        u, c = pred.chunk(2)
        v = sde.cfg_combine(u, c, guidance_scale)                       # fast.py:640-642 (bf16 ops when bf16_io)
        noise = None if noises is None else noises[i]
        prev, lp, _, _ = sde.sde_step_with_logprob_new(sch.sigmas, [i], v.float(), latents.float(), cur, noise=noise)
        latents = cast(prev)                                            # fast.py:654-655
        if in_window:
            all_latents.append(latents)
            all_log_probs.append(lp)
            all_timesteps.append(t.repeat(len(latents)))
    image = vae.decode_latents_to_image(vae_params, latents) if decode else None
    return image, all_latents, all_log_probs, all_timesteps, sch


def replay_loss(mmdit, sch, latents, next_latents, step_index, embeds, pooled_all, old_log_prob, advantages,
                guidance_scale, noise_level, clip_range, adv_clip_max, bf16_io=True):
    """train_pick:233-267 + :1111-1123 for one (batch, j)."""
    t = sch.timesteps[step_index].expand(2 * latents.shape[0])
    pred = mmdit.forward(torch.cat([latents] * 2).float(), t, embeds.float(), pooled_all.float())
    u, c = pred.chunk(2)
    v = u + guidance_scale * (c - u)
    _, lp, _, _ = sde.sde_step_with_logprob_new(sch.sigmas, [step_index] * latents.shape[0], v, latents.float(),
                                                noise_level, prev_sample=next_latents.float())
    loss, info = grpo_loss.grpo_clip_loss(lp, old_log_prob, advantages, clip_range, adv_clip_max)
    return loss, lp, info
