"""Drop-in for `pipeline_with_logprob_random` of
`adv_grpo/diffusers_patch/sd3_pipeline_with_logprob_fast.py:453-674` (the only rollout function the two
training scripts import, train_sd3_fast_pickscore.py:20): same keyword surface and the same
`(image, all_latents, all_log_probs, all_timesteps)` return.

What changes underneath: one MMDiT forward per step on the CFG batch [negative ; positive] (replayed
from a CUDA graph when the pipeline provides one), then ONE fused kernel per step for
CFG-combine + Flow-CPS step + log-prob + bf16 cast of the next latents (the reference launches ~20
kernels and synchronises with the host 2-9 times per step in `index_for_timestep().item()`).
Quirks kept: the SDE window is [random_timestep, random_timestep + train_num_steps); outside it the
same step runs with noise_level 0 (still consuming an RNG draw, Q2); stored latents are bf16 (Q4).
"""
import random

import torch

from .. import ops
from ..scheduler import retrieve_timesteps
from .sd3_sde_with_logprob import _next_offset


@torch.no_grad()
def pipeline_with_logprob_random(self, prompt=None, prompt_2=None, prompt_3=None, height=None, width=None,
                                 num_inference_steps=28, mini_num_image_per_prompt=1, sigmas=None,
                                 guidance_scale=7.0, negative_prompt=None, negative_prompt_2=None,
                                 negative_prompt_3=None, generator=None, latents=None, prompt_embeds=None,
                                 negative_prompt_embeds=None, pooled_prompt_embeds=None,
                                 negative_pooled_prompt_embeds=None, output_type="pil",
                                 joint_attention_kwargs=None, clip_skip=None,
                                 callback_on_step_end_tensor_inputs=("latents",), max_sequence_length=256,
                                 skip_layer_guidance_scale=2.8, noise_level=0.7, train_num_steps=1,
                                 process_index=0, sample_num_steps=10, random_timestep=None, noise=None):
    if prompt_embeds is None or pooled_prompt_embeds is None:
        raise ValueError("pass prompt_embeds / pooled_prompt_embeds (the training scripts always do, "
                         "train_sd3_fast_pickscore.py:755-772); text encoding is outside this path")
    height = height or self.default_sample_size * self.vae_scale_factor
    width = width or self.default_sample_size * self.vae_scale_factor
    device = self._execution_device
    do_cfg = guidance_scale > 1
    self._guidance_scale = guidance_scale
    G = mini_num_image_per_prompt
    prompt_embeds = prompt_embeds.repeat(G, 1, 1)                                    # fast.py:551-554
    pooled_prompt_embeds = pooled_prompt_embeds.repeat(G, 1)
    if do_cfg:
        negative_prompt_embeds = negative_prompt_embeds.repeat(G, 1, 1)
        negative_pooled_prompt_embeds = negative_pooled_prompt_embeds.repeat(G, 1)
    B = prompt_embeds.shape[0]
    C = self.transformer.config.in_channels
    if latents is None:                                                              # prepare_latents, fast.py:559-568
        latents = torch.randn((B, C, height // self.vae_scale_factor, width // self.vae_scale_factor),
                              generator=generator, device=device, dtype=torch.float32).to(prompt_embeds.dtype)
    latents = latents.to(device=device, dtype=torch.bfloat16).contiguous()
    timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, sigmas=sigmas)
    random.seed(process_index)                                                       # fast.py:585-587
    if random_timestep is None:
        random_timestep = random.randint(0, sample_num_steps // 2)
    if do_cfg:
        embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
        pooled = torch.cat([negative_pooled_prompt_embeds, pooled_prompt_embeds], dim=0)
    else:
        embeds, pooled = prompt_embeds, pooled_prompt_embeds
    seed = generator.initial_seed() if generator is not None else torch.initial_seed()
    all_latents, all_log_probs, all_timesteps = [], [], []
    transformer = getattr(self, "graphed_transformer", None) or self.transformer
    for i in range(len(timesteps)):
        t = timesteps[i:i + 1]
        in_window = random_timestep <= i < random_timestep + train_num_steps
        cur_noise_level = noise_level if in_window else 0                            # fast.py:606-623
        if i == random_timestep:
            all_latents.append(latents)
        model_in = torch.cat([latents, latents]) if do_cfg else latents
        noise_pred = transformer(hidden_states=model_in, timestep=t.expand(model_in.shape[0]),
                                 encoder_hidden_states=embeds, pooled_projections=pooled, return_dict=False)[0]
        if do_cfg:
            vu, vt = noise_pred[:B], noise_pred[B:]
        else:
            vu, vt = None, noise_pred
        step_noise = None if noise is None else noise[i]
        latents, log_prob, _, _ = ops.cfg_sde_step_logprob(
            vu, vt, latents, t, self.scheduler.timesteps, self.scheduler.sigmas, guidance_scale, cur_noise_level,
            noise=step_noise, seed=seed, offset=_next_offset(latents.numel()))
        if in_window:
            all_latents.append(latents)
            all_log_probs.append(log_prob)
            all_timesteps.append(t.repeat(B))
    z = latents.float() / self.vae.config.scaling_factor + self.vae.config.shift_factor   # fast.py:667
    image = self.vae.decode(z.to(self.vae.dtype), return_dict=False)[0]
    image = self.image_processor.postprocess(image, output_type=output_type)
    return image, all_latents, all_log_probs, all_timesteps
