"""Wall-clock (synchronised) time of each phase of one GRPO step at BASELINE config 2 on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import weights
from adv_grpo_b200.config import load_config
from adv_grpo_b200.diffusers_patch.sd3_pipeline_with_logprob_fast import pipeline_with_logprob_random
from adv_grpo_b200.pickscore_scorer import PickScoreScorer
from adv_grpo_b200.pipeline import StableDiffusion3Pipeline
from adv_grpo_b200.trainer import GRPOTrainer

dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
pipe = StableDiffusion3Pipeline.from_seed(weights.SD35_MEDIUM, weights.VAE_SD3, device=dev, seed=0)
scorer = PickScoreScorer(device=dev, dtype=torch.bfloat16)
cfg = load_config("pickscore_cotrain_sd3_fast")
cfg.sample.num_batches_per_epoch = 2
cfg.train.gradient_accumulation_steps = 1
cfg.train_d = False
tr = GRPOTrainer(cfg, pipe, [f"synthetic prompt {i}" for i in range(9)], scorer=scorer, device=dev)
for _ in range(3):
    tr.run_epoch()

def timed(fn, *a):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(*a); torch.cuda.synchronize()
    return r, (time.perf_counter() - t) * 1e3

for rep in range(2):
    samples, t_s = timed(tr.sample_epoch)
    adv, t_a = timed(tr.compute_advantages, samples)
    _, t_t = timed(tr.train_generator, samples, adv)
    print(f"sample_epoch {t_s:.1f} ms | advantages {t_a:.1f} ms | train_generator {t_t:.1f} ms | total {t_s + t_a + t_t:.1f} ms")
# inside sampling
pe, pp = tr.embedder(0)
def roll():
    return pipeline_with_logprob_random(pipe, prompt_embeds=pe, pooled_prompt_embeds=pp, negative_prompt_embeds=tr.neg_embeds,
        negative_pooled_prompt_embeds=tr.neg_pooled, num_inference_steps=10, guidance_scale=4.5, output_type="pt", height=512, width=512,
        noise_level=0.8, mini_num_image_per_prompt=8, train_num_steps=2, process_index=0, sample_num_steps=10, random_timestep=0,
        generator=tr.generator)
(images, lats, lps, tss), t_r = timed(roll)
lat = lats[-1]
_, t_v = timed(lambda: pipe.vae.decode(lat.float() / 1.5305 + 0.0609))
_, t_sc = timed(lambda: tr._score(images, ["p"] * 8))
x = torch.cat([lat, lat]); t = tss[0].repeat(2); e = torch.cat([tr.neg_embeds.repeat(8, 1, 1), pe.repeat(8, 1, 1)]); po = torch.cat([tr.neg_pooled.repeat(8, 1), pp.repeat(8, 1)])
with torch.no_grad():
    _, t_f = timed(lambda: pipe.graphed_transformer(x, t, e, po))
    _, t_fe = timed(lambda: pipe.transformer(x, t, e, po))
print(f"rollout(10 steps + vae) {t_r:.1f} ms | vae decode {t_v:.1f} ms | pickscore(8 imgs) {t_sc:.1f} ms | mmdit fwd graph {t_f:.1f} ms eager {t_fe:.1f} ms")
def opt():
    tr.optimizer.step(); tr.transformer.invalidate_lora_cache()
    with torch.no_grad():
        tr.transformer._pack_lora()
_, t_o = timed(opt)
print(f"clip+adamw+zero+repack {t_o:.1f} ms")
