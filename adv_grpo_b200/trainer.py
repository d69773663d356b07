"""The GRPO epoch of `scripts/train_sd3_fast_pickscore.py:709-1191` / `train_sd3_fast_dino_patch.py`
on the B200 kernels: SAMPLING (rollout + reward) -> gather + group advantage -> (discriminator step |
generator step with the clipped policy-gradient loss).  One process per GPU; prompts shard across
ranks (each rank rolls out whole G-sample groups); NCCL is used only for
  * the packed all-gather of rewards + prompt ids (`train_pick:926-938,966`),
  * the flat LoRA-gradient all-reduce at each accumulation boundary (`:1165-1171`, ZeRO-2 there),
  * the discriminator parameter sync after a D step (north_star; the reference leaves per-rank
    discriminators un-synchronised, quirk Q3 -- `sync_discriminator=False` reproduces that).
Everything between the collectives stays on the device: no `.cpu().numpy()` of rewards, no tokenizer
decode of prompt ids, no float64 host advantages.
"""
import os
import warnings

import torch
import torch.distributed as dist

from . import ops
from .diffusers_patch.sd3_pipeline_with_logprob_fast import pipeline_with_logprob_random
from .dinov2 import dino_hinge_d_loss
from .ema import EMAModuleWrapper
from .optim import FlatClipAdamW, TorchOrderAdam
from .pick_score_training import CLIPCriterion, CLIPCriterionConfig
from .pickscore_scorer import images_to_pixel_values
from .rewards import multi_score
from .sampler import DistributedKRepeatSampler
from .stat_tracking import PerPromptStatTracker


class SyntheticTextEmbedder:
    """Stand-in for `compute_text_embeddings` (`train_pick:186-193`; the CLIP-L/G + T5-XXL encoders are
    outside the hot path and no weights exist here): deterministic N(0,1) embeddings per prompt index
    (SURVEY.md section 8d)."""

    def __init__(self, joint_dim=4096, pooled_dim=2048, n_tokens=205, device="cuda"):
        self.joint_dim, self.pooled_dim, self.n_tokens, self.device = joint_dim, pooled_dim, n_tokens, device

    def __call__(self, prompt_index):
        g = torch.Generator().manual_seed(1000 + int(prompt_index))
        e = torch.randn(1, self.n_tokens, self.joint_dim, generator=g).to(self.device, torch.bfloat16)
        p = torch.randn(1, self.pooled_dim, generator=g).to(self.device, torch.bfloat16)
        return e, p

    def negative(self):
        g = torch.Generator().manual_seed(999)
        e = torch.randn(1, self.n_tokens, self.joint_dim, generator=g).to(self.device, torch.bfloat16)
        p = torch.randn(1, self.pooled_dim, generator=g).to(self.device, torch.bfloat16)
        return e, p


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_cat(t):
    rank, world = _world()
    if world == 1:
        return t
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous())
    return out


def compute_log_prob(transformer, pipeline, sample, j, embeds, pooled_embeds, config, mean_ref=None, want_mean=False):
    """`train_pick:233-267`: transformer forward on the CFG batch + fused CFG / Flow-CPS replay log-prob.
    With `mean_ref` (the adapter-disabled prev_sample_mean, `train_pick:1105-1108`) the last element of the
    returned tuple is the per-sample KL term instead of std_dev_t's successor (5-tuple)."""
    lat = sample["latents"][:, j]
    ts = sample["timesteps"][:, j]
    cfg = bool(config.train.cfg)
    if cfg:
        noise_pred = transformer(hidden_states=torch.cat([lat, lat]), timestep=torch.cat([ts, ts]),
                                 encoder_hidden_states=embeds, pooled_projections=pooled_embeds, return_dict=False)[0]
    else:
        noise_pred = transformer(hidden_states=lat, timestep=ts, encoder_hidden_states=embeds,
                                 pooled_projections=pooled_embeds, return_dict=False)[0]
    out = ops.sde_logprob_replay(noise_pred, lat, sample["next_latents"][:, j], ts,
                                 pipeline.scheduler.timesteps, pipeline.scheduler.sigmas,
                                 config.sample.guidance_scale, config.sample.noise_level, cfg=cfg,
                                 want_mean=want_mean, mean_ref=mean_ref)
    if mean_ref is not None:
        log_prob, mean, std, kl = out
        return sample["next_latents"][:, j], log_prob, mean, std, kl
    log_prob, mean, std = out
    return sample["next_latents"][:, j], log_prob, mean, std


class GraphedMicroStep:
    """One replay micro-step (MMDiT forward on the CFG batch + fused CFG/SDE log-prob + clipped GRPO loss +
    backward into the LoRA .grad buffers) captured as a CUDA graph: the eager version is CPU-bound (~3000
    launches of Python/autograd/ctypes dispatch per micro-step).  First call per shape runs eagerly (warm-up),
    the second captures, later ones replay.  Gradients accumulate in place into the existing .grad tensors."""

    def __init__(self, trainer):
        self.t, self.entries = trainer, {}

    def _eager(self, lat, nxt, ts, embeds, pooled, old_lp, adv, sched_t, sigmas, grad_scale):
        tr, c = self.t, self.t.config
        noise_pred = tr.transformer(hidden_states=torch.cat([lat, lat]), timestep=torch.cat([ts, ts]),
                                    encoder_hidden_states=embeds, pooled_projections=pooled, return_dict=False)[0]
        # the scheduler tables are re-created by every rollout: they enter the graph through static copies
        log_prob, _, _ = ops.sde_logprob_replay(noise_pred, lat, nxt, ts, sched_t, sigmas, c.sample.guidance_scale,
                                                c.sample.noise_level, cfg=True)
        loss, stats = ops.grpo_clip_loss(log_prob, old_lp, adv, c.train.clip_range, c.train.adv_clip_max,
                                         grad_scale=grad_scale)
        loss.backward()
        return stats

    def __call__(self, lat, nxt, ts, embeds, pooled, old_lp, adv, grad_scale):
        from . import _lib
        key = (tuple(lat.shape), tuple(embeds.shape), float(grad_scale))
        ent = self.entries.get(key)
        sch = self.t.pipeline.scheduler
        args = (lat, nxt, ts, embeds, pooled, old_lp, adv, sch.timesteps.to(lat.device, torch.float32).contiguous(),
                sch.sigmas.to(lat.device, torch.float32).contiguous())
        if ent is None:                                   # warm-up (also creates the .grad tensors)
            self.entries[key] = "warm"
            return self._eager(*args, grad_scale).clone()
        if ent == "warm":
            static = [a.clone() for a in args]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g):
                stats = self._eager(*static, grad_scale)
            ent = (g, static, stats, _lib.launch_count() - n0)
            self.entries[key] = ent
        g, static, stats, n_kernels = ent
        for dst, src in zip(static, args):
            dst.copy_(src)
        g.replay()
        _lib.add_launches(n_kernels)
        return stats.clone()


class GRPOTrainer:
    def __init__(self, config, pipeline, prompts, scorer=None, head=None, embedder=None, device="cuda",
                 reference_image_fn=None, sync_discriminator=True, graph_train=None):
        self.config, self.pipeline, self.prompts, self.device = config, pipeline, list(prompts), device
        self.rank, self.world = _world()
        self.transformer = pipeline.transformer
        self.scorer, self.head = scorer, head
        self.embedder = embedder or SyntheticTextEmbedder(self.transformer.cfg["joint_dim"],
                                                          self.transformer.cfg["pooled_dim"], device=device)
        self.reference_image_fn = reference_image_fn or self._synthetic_reference
        self.sync_discriminator = sync_discriminator
        s, t = config.sample, config.train
        if not config.get("use_lora", True):                               # train_pick:488: every transformer weight trains
            if t.beta > 0:
                raise NotImplementedError("train.beta > 0 needs the adapter-disabled reference forward (train_pick:1105-1108), "
                                          "which only exists with use_lora=True")
            self.transformer.enable_full_finetune()
            graph_train = False                                            # the full-gradient replay runs eagerly
        elif t.get("lora_path", None):                                     # resume, train_pick:506-509
            self.transformer.load_adapter(t.lora_path)
        self.params = self.transformer.trainable_parameters()
        # clip_grad_norm_ + AdamW.step + zero_grad of train_pick:1165-1171 as one native call on the flat parameter
        self.optimizer = FlatClipAdamW(self.params, lr=t.learning_rate, betas=(t.adam_beta1, t.adam_beta2),
                                       weight_decay=t.adam_weight_decay, eps=t.adam_epsilon,
                                       max_grad_norm=t.max_grad_norm)
        self.ema = EMAModuleWrapper(self.params, decay=0.9, update_step_interval=8, device=device) if t.ema else None
        self.reward_fn = multi_score(device, dict(config.reward_fn))
        self.reward_key = next(iter(dict(config.reward_fn)))
        # k = num_image_per_prompt // mini_num_image_per_prompt as in the reference (train_pick:577): a prompt is rolled
        # out on k ranks and its advantage group has k * mini images.  Only when world * train_batch_size is not divisible
        # by k (single-GPU runs of the 8-GPU presets) does the group collapse to one rank-call, with a warning.
        # `sample.shard_groups_across_ranks = False` forces k = 1 (north_star bench: every rank owns whole groups).
        k_ref = max(1, s.num_image_per_prompt // s.mini_num_image_per_prompt)
        k = k_ref
        if s.get("shard_groups_across_ranks", None) is False:
            k = 1
        elif (self.world * s.train_batch_size) % k_ref != 0:
            k = 1
            warnings.warn(f"world_size * train_batch_size = {self.world * s.train_batch_size} is not divisible by "
                          f"k = num_image_per_prompt // mini_num_image_per_prompt = {k_ref}: every prompt is rolled out "
                          f"by one rank-call only (group size {s.mini_num_image_per_prompt}, not {s.num_image_per_prompt})")
        self.sampler = DistributedKRepeatSampler(self.prompts, s.train_batch_size, k, self.world, self.rank, seed=42)
        self.stat_tracker = PerPromptStatTracker(s.global_std, device=device)
        self.neg_embeds, self.neg_pooled = self.embedder.negative()
        self.generator = torch.Generator(device=device).manual_seed(config.seed + self.rank)
        self.global_step = 0
        self.epoch = 0
        self._micro = 0                      # replay micro-steps so far (accumulation counter, persists across epochs)
        self.d_schedule = None               # optional epoch -> bool override of the D / G gate (bench: alternate)
        self.optimizer_D = None
        if config.get("train_d", False):
            if self.reward_key == "pickscore_cotrain":
                self.criterion = CLIPCriterion(CLIPCriterionConfig())
                # Adam(lr = d_lr, betas = (0.5, 0.999)) of train_pick:658 / train_dino:750 in torch's op order on one native
                # pass per tensor (bf16 parameters keep the reference's bf16 update arithmetic)
                self.optimizer_D = TorchOrderAdam(scorer.model.parameters(), lr=config.d_lr, betas=(0.5, 0.999))
            elif head is not None:
                self.optimizer_D = TorchOrderAdam(head.parameters(), lr=config.d_lr, betas=(0.5, 0.999))
        self.last_info = {}
        if graph_train is None:
            graph_train = getattr(pipeline, "graphed_transformer", None) is not None
        self.micro_step = GraphedMicroStep(self) if graph_train else None

    # ------------------------------------------------------------------ sampling
    def _synthetic_reference(self, prompt_index, n, size):
        g = torch.Generator().manual_seed(11 + int(prompt_index))
        return torch.rand(n, 3, size, size, generator=g).to(self.device)

    def _prompt_ids(self, prompt, n):
        ids = self.pipeline.tokenizer([prompt], padding="max_length", max_length=256, truncation=True)["input_ids"]
        return ids.to(self.device).repeat(n, 1)

    def _score(self, images, prompts):
        kw = dict(scorer=self.scorer)
        if self.head is not None:
            kw["head"] = self.head
        details, _ = self.reward_fn(images.to(torch.bfloat16), prompts, [{}] * len(prompts), **kw)
        return {k: torch.as_tensor(v, device=self.device).float() for k, v in details.items()}

    @torch.no_grad()
    def sample_epoch(self):
        c, s = self.config, self.config.sample
        G = s.mini_num_image_per_prompt
        samples = []
        self.transformer.eval()
        for i in range(s.num_batches_per_epoch):
            idx = self.sampler.indices_for_epoch(self.epoch * s.num_batches_per_epoch + i)[self.rank][0]
            prompt = self.prompts[idx]
            pe, pp = self.embedder(idx)
            images, latents, log_probs, timesteps = pipeline_with_logprob_random(
                self.pipeline, prompt_embeds=pe, pooled_prompt_embeds=pp, negative_prompt_embeds=self.neg_embeds,
                negative_pooled_prompt_embeds=self.neg_pooled, num_inference_steps=s.num_steps,
                guidance_scale=s.guidance_scale, output_type="pt", height=c.resolution, width=c.resolution,
                noise_level=s.noise_level, mini_num_image_per_prompt=G, train_num_steps=s.train_num_steps,
                process_index=self.rank, sample_num_steps=s.num_steps, random_timestep=s.get("random_timestep", 0),
                generator=self.generator)
            ref_images = self.reference_image_fn(idx, G, c.resolution)
            prompts = [prompt] * G
            lat = torch.stack(latents, dim=1)                       # [G, T+1, 16, h, w]   train_pick:806-810
            # generated and reference images through the reward model as ONE batch of 2G (the reference submits two
            # reward_fn calls, train_pick:816-817; scores are per image, so the split halves are the same values)
            both = self._score(torch.cat([images.to(torch.float32), ref_images.to(images.device, torch.float32)]),
                               prompts + prompts)
            samples.append({
                "prompt_ids": self._prompt_ids(prompt, G),
                "prompt_embeds": pe.repeat(G, 1, 1), "pooled_prompt_embeds": pp.repeat(G, 1),
                "timesteps": torch.stack(timesteps, dim=1), "latents": lat[:, :-1], "next_latents": lat[:, 1:],
                "log_probs": torch.stack(log_probs, dim=1),
                "rewards": {k: v[:G] for k, v in both.items()}, "reference_rewards": {k: v[G:] for k, v in both.items()},
                "images": images, "ref_images": ref_images, "prompts": prompts,
            })
        return samples

    # ------------------------------------------------------------------ advantages
    def compute_advantages(self, samples):
        T = self.config.sample.train_num_steps
        avg = torch.cat([s["rewards"]["avg"] for s in samples])                 # [N_local]
        rewards = avg[:, None].repeat(1, T)                                       # train_pick:928
        ids = torch.cat([s["prompt_ids"] for s in samples])
        g_rewards, g_ids = all_gather_cat(rewards), all_gather_cat(ids)           # train_pick:930,966
        adv = self.stat_tracker.update_device(g_ids, g_rewards)                   # float64 [N_total, T]
        zero_std_ratio, reward_std_mean = self.stat_tracker.zero_std_stats()
        group_size, _ = self.stat_tracker.get_stats()
        self.stat_tracker.clear()
        n = rewards.shape[0]
        local = adv.reshape(self.world, n, T)[self.rank]                          # train_pick:995-999
        self.last_info.update(reward_mean=g_rewards[:, 0].mean(), zero_std_ratio=zero_std_ratio,
                              reward_std_mean=reward_std_mean, group_size=group_size)
        return local

    # ------------------------------------------------------------------ discriminator step
    def discriminator_step(self, samples):
        c = self.config
        real = torch.cat([s["ref_images"] for s in samples])
        fake = torch.cat([s["images"] for s in samples])
        prompts = [p for s in samples for p in s["prompts"]]
        if self.reward_key == "pickscore_cotrain":
            model = self.scorer.model
            for p in model.parameters():
                p.requires_grad = False
            for p in model.vision_model.encoder.layers[c.tune_layer:].parameters():   # train_pick:1016-1020
                p.requires_grad = True
            ids = self.scorer.processor.tokenizer(prompts, padding="max_length", truncation=True, max_length=77)["input_ids"].to(self.device)
            # tensor_to_pil_list (train_pick:133-148): (x*255).astype(uint8) truncation, then CLIPProcessor
            to_u8 = lambda x: (x.float().clamp(0, 1) * 255).to(torch.uint8)
            batch = {"input_ids": ids, "pixels_0": images_to_pixel_values(to_u8(real), self.device),
                     "pixels_1": images_to_pixel_values(to_u8(fake), self.device),
                     "label_0": torch.tensor(1.0, device=self.device), "label_1": torch.tensor(0.0, device=self.device),
                     "num_examples_per_prompt": torch.tensor(1.0, device=self.device)}
            loss = self.criterion(model, batch)
            params = [p for p in model.parameters() if p.requires_grad]
        else:
            with torch.no_grad():
                fr = self.scorer.forward_features(ops.dino_preprocess(real, 518))
                ff = self.scorer.forward_features(ops.dino_preprocess(fake, 518))
            N = fr.shape[1] - 1
            n_sel = min(64, N)
            ir = torch.randint(0, N, (fr.shape[0], n_sel), device=self.device)
            if_ = torch.randint(0, N, (ff.shape[0], n_sel), device=self.device)
            loss, acc = dino_hinge_d_loss(self.head, fr, ff, ir, if_, 0.3)               # train_dino:186-219
            self.last_info["d_acc"] = acc.detach()
            params = list(self.head.parameters())
        trace = os.environ.get("ADVGRPO_TRACE_DSTEP", "0") == "1"      # CUDA-event split of the D step into last_info
        if trace:
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            evs[0].record()
        self.optimizer_D.zero_grad()
        loss.backward()
        if trace:
            evs[1].record()
        if self.world > 1 and self.sync_discriminator:
            flat = torch.cat([p.grad.reshape(-1).float() for p in params])
            dist.all_reduce(flat)
            flat /= self.world
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()
        if trace:
            evs[2].record()
        self.optimizer_D.step()
        if trace:
            evs[3].record()
            torch.cuda.synchronize()
            self.last_info["d_step_ms"] = dict(backward=evs[0].elapsed_time(evs[1]), grad_sync=evs[1].elapsed_time(evs[2]),
                                               optimizer=evs[2].elapsed_time(evs[3]))
        return loss.detach()

    # ------------------------------------------------------------------ generator step
    def _sync_grads(self):
        if self.world == 1:
            return
        grads = [p.grad for p in self.params]
        if len(grads) == 1:                          # the flat LoRA master parameter: reduce its gradient in place
            dist.all_reduce(grads[0])
            grads[0] /= self.world
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        flat /= self.world
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def train_generator(self, samples, advantages):
        """`train_pick:1050-1180`: per sample batch and SDE-window step one replay micro-step (forward + loss +
        backward), an optimizer step every `gradient_accumulation_steps * T` micro-steps (the accelerate accumulation
        counter persists across epochs), `train.num_inner_epochs` passes over the epoch's samples.  `train.micro_batch`
        (not in the reference) replays a group in chunks of that many samples (same gradient: the loss is a batch mean,
        each chunk is scaled by its share) to bound activation memory at 1024x1024 / G = 16."""
        c, t = self.config, self.config.train
        T = c.sample.train_num_steps
        gas = t.gradient_accumulation_steps * T                      # Accelerator(grad_accum = gas * T), train_pick:426
        self.transformer.train()
        stats_acc = []
        n_local = samples[0]["latents"].shape[0]
        mb = int(t.get("micro_batch", 0) or 0)
        mb = mb if 0 < mb < n_local and n_local % mb == 0 else n_local
        for _inner in range(int(t.get("num_inner_epochs", 1) or 1)):
            for i, sample in enumerate(samples):
                adv_i = advantages[i * n_local:(i + 1) * n_local]
                for j in range(T):
                    stats_j = None
                    for c0 in range(0, n_local, mb):
                        sl = slice(c0, c0 + mb)
                        sub = sample if mb == n_local else {k: sample[k][sl] for k in ("latents", "next_latents", "timesteps",
                                                                                       "log_probs", "prompt_embeds",
                                                                                       "pooled_prompt_embeds")}
                        if t.cfg:
                            embeds = torch.cat([self.neg_embeds.repeat(mb, 1, 1), sub["prompt_embeds"]])
                            pooled = torch.cat([self.neg_pooled.repeat(mb, 1), sub["pooled_prompt_embeds"]])
                        else:
                            embeds, pooled = sub["prompt_embeds"], sub["pooled_prompt_embeds"]
                        adv_c = adv_i[sl, j].contiguous()
                        share = mb / n_local
                        if t.beta > 0:
                            # KL-regularised step (train_pick:1105-1108,1124-1128): reference mean from the adapter-disabled
                            # forward, loss = policy_loss + beta * mean_b(mean_chw((mu - mu_ref)^2))
                            with torch.no_grad(), self.transformer.disable_adapter():
                                _, _, mean_ref, _ = compute_log_prob(self.transformer, self.pipeline, sub, j, embeds, pooled, c,
                                                                     want_mean=True)
                            _, log_prob, _, _, kl = compute_log_prob(self.transformer, self.pipeline, sub, j, embeds, pooled, c,
                                                                     mean_ref=mean_ref)
                            loss, stats = ops.grpo_clip_loss(log_prob, sub["log_probs"][:, j], adv_c,
                                                             t.clip_range, t.adv_clip_max, grad_scale=share / gas)
                            kl_loss = kl.mean()
                            (loss + (t.beta * share / gas) * kl_loss.to(loss.dtype)).backward()
                            self.last_info["kl_loss"] = kl_loss.detach()
                            stats = stats.clone()
                            stats[0] = stats[5] + t.beta * kl_loss.to(stats.dtype)       # logged loss = policy + beta * kl
                        elif self.micro_step is not None and t.cfg:
                            stats = self.micro_step(sub["latents"][:, j].contiguous(), sub["next_latents"][:, j].contiguous(),
                                                    sub["timesteps"][:, j].contiguous(), embeds, pooled,
                                                    sub["log_probs"][:, j].contiguous(), adv_c, share / gas)
                        else:
                            _, log_prob, _, _ = compute_log_prob(self.transformer, self.pipeline, sub, j, embeds, pooled, c)
                            loss, stats = ops.grpo_clip_loss(log_prob, sub["log_probs"][:, j], adv_c,
                                                             t.clip_range, t.adv_clip_max, grad_scale=share / gas)
                            loss.backward()
                        stats_j = stats * share if stats_j is None else stats_j + stats * share
                    stats_acc.append(stats_j)
                    self._micro += 1
                    if self._micro % gas == 0:
                        self._sync_grads()
                        self.optimizer.step()                       # clip + AdamW + gradient clear (csrc/optim.cu)
                        self.transformer.invalidate_lora_cache()
                        self.global_step += 1
                if self.ema is not None:
                    self.ema.step(self.params, self.global_step)
        if stats_acc:
            m = torch.stack(stats_acc).mean(0)
            self.last_info.update(loss=m[0], approx_kl=m[1], clipfrac=m[2], clipfrac_gt_one=m[3], clipfrac_lt_one=m[4],
                                  policy_loss=m[5])

    # ------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def evaluate(self, prompt_indices=None, batch_size=None):
        """`eval` of `train_pick:269-382` (every `config.eval_freq` epochs): EMA weights swapped in, the test prompts in
        batches of `sample.test_batch_size` through the SAME rollout function as a deterministic ODE (`noise_level=0`,
        `sample.eval_num_steps` steps, one image per prompt, generator seeded with 0 per batch), rewards gathered over
        ranks, `eval_reward_<key>` = mean over the values != -10.  Returns (metrics, last_batch_images); the wandb image
        logging of the reference is not part of this path."""
        c, s = self.config, self.config.sample
        idxs = list(range(len(self.prompts))) if prompt_indices is None else list(prompt_indices)
        bs = int(batch_size or s.test_batch_size)
        # test DataLoader sharded by accelerate, which pads the last batches so that every rank runs the same number of
        # equally sized batches (the per-batch gathers below would otherwise mismatch): pad with repeats, mask them out
        n_valid = len(idxs)
        per_rank = -(-max(n_valid, 1) // (self.world * bs)) * bs
        padded = idxs + [idxs[i % n_valid] for i in range(per_rank * self.world - n_valid)] if n_valid else []
        valid_flags = [1.0] * n_valid + [0.0] * (len(padded) - n_valid)
        idxs, flags = padded[self.rank::self.world], valid_flags[self.rank::self.world]
        if c.train.ema and self.ema is not None:
            self.ema.copy_ema_to(self.params, store_temp=True)
            self.transformer.invalidate_lora_cache()
        self.transformer.eval()
        all_rewards, images = {}, None
        try:
            for b0 in range(0, len(idxs), bs):
                batch = idxs[b0:b0 + bs]
                prompts = [self.prompts[i] for i in batch]
                emb = [self.embedder(i) for i in batch]
                pe, pp = torch.cat([e[0] for e in emb]), torch.cat([e[1] for e in emb])
                generator = torch.Generator(device=self.device).manual_seed(0)
                images, _, _, _ = pipeline_with_logprob_random(
                    self.pipeline, prompt_embeds=pe, pooled_prompt_embeds=pp,
                    negative_prompt_embeds=self.neg_embeds.repeat(len(batch), 1, 1),
                    negative_pooled_prompt_embeds=self.neg_pooled.repeat(len(batch), 1),
                    num_inference_steps=s.eval_num_steps, guidance_scale=s.guidance_scale, output_type="pt",
                    height=c.resolution, width=c.resolution, noise_level=0, mini_num_image_per_prompt=1,
                    process_index=self.rank, sample_num_steps=s.num_steps, random_timestep=s.get("random_timestep", 0),
                    generator=generator)
                keep_b = all_gather_cat(torch.tensor(flags[b0:b0 + bs], device=self.device)) > 0
                for k, v in self._score(images, prompts).items():
                    all_rewards.setdefault(k, []).append(all_gather_cat(v)[keep_b])
        finally:
            if c.train.ema and self.ema is not None:
                self.ema.copy_temp_to(self.params)
                self.transformer.invalidate_lora_cache()
        metrics = {}
        for k, vals in all_rewards.items():
            v = torch.cat(vals)
            keep = v != -10
            metrics[f"eval_reward_{k}"] = v[keep].mean() if keep.any() else v.new_tensor(float("nan"))
        return metrics, images

    # ------------------------------------------------------------------ checkpoint
    def save_ckpt(self, save_dir):
        """`save_ckpt` (`train_pick:389-398`): peft adapter directory with the EMA weights swapped in, rank 0 only."""
        from .checkpoint import save_ckpt
        return save_ckpt(save_dir, self.transformer, self.global_step, ema=self.ema, trainable_parameters=self.params,
                         use_ema=bool(self.config.train.ema), is_main_process=self.rank == 0)

    # ------------------------------------------------------------------ one epoch
    def run_epoch(self):
        c = self.config
        samples = self.sample_epoch()
        advantages = self.compute_advantages(samples)
        did_d = False
        if c.get("train_d", False) and self.optimizer_D is not None:
            if self.d_schedule is not None:
                did_d = bool(self.d_schedule(self.epoch))
            elif self.reward_key == "pickscore_cotrain":
                gen = all_gather_cat(torch.cat([s["rewards"][self.reward_key] for s in samples])).mean()
                ref = all_gather_cat(torch.cat([s["reference_rewards"][self.reward_key] for s in samples])).mean()
                did_d = bool(ref < gen)                                        # train_pick:1025 (one host sync per epoch)
            else:
                did_d = (self.epoch + 1) % c.d_times != 0                      # train_dino:1097
        if did_d:
            self.last_info["d_loss"] = self.discriminator_step(samples)
            self.global_step += 1                                              # quirk Q8: G step skipped this epoch
        else:
            self.train_generator(samples, advantages)
        self.epoch += 1
        return {"did_d_step": did_d, "n_samples": sum(s["latents"].shape[0] for s in samples), **self.last_info}
