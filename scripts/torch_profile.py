"""torch.profiler view (op names + input shapes + python stack) of one training micro-step and one rollout step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from adv_grpo_b200 import weights
from adv_grpo_b200.config import load_config
from adv_grpo_b200.pickscore_scorer import PickScoreScorer
from adv_grpo_b200.pipeline import StableDiffusion3Pipeline
from adv_grpo_b200.trainer import GRPOTrainer

dev = "cuda:0"
pipe = StableDiffusion3Pipeline.from_seed(weights.SD35_MEDIUM, weights.VAE_SD3, device=dev, seed=0, use_cuda_graph=False)
scorer = PickScoreScorer(device=dev, dtype=torch.bfloat16)
cfg = load_config("pickscore_cotrain_sd3_fast")
cfg.sample.num_batches_per_epoch = 1
cfg.train.gradient_accumulation_steps = 1
cfg.train_d = False
tr = GRPOTrainer(cfg, pipe, [f"synthetic prompt {i}" for i in range(9)], scorer=scorer, device=dev)
tr.run_epoch()
samples = tr.sample_epoch()
adv = tr.compute_advantages(samples)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    tr.train_generator(samples, adv)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
print(prof.key_averages(group_by_stack_n=4).table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=50))
