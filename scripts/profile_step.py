"""One GRPO group (rollout 10 steps + VAE + PickScore x2 + advantage + 2 replay micro-steps + optimizer step)
at BASELINE config 2 between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from adv_grpo_b200 import weights
from adv_grpo_b200.config import load_config
from adv_grpo_b200.pickscore_scorer import PickScoreScorer
from adv_grpo_b200.pipeline import StableDiffusion3Pipeline
from adv_grpo_b200.trainer import GRPOTrainer

dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = True        # as bench.py / train_sd3_fast_pickscore.py:537-538
torch.backends.cudnn.allow_tf32 = True
graph = os.environ.get("GRAPH", "0") == "1"
pipe = StableDiffusion3Pipeline.from_seed(weights.SD35_MEDIUM, weights.VAE_SD3, device=dev, seed=0, use_cuda_graph=graph)
scorer = PickScoreScorer(device=dev, dtype=torch.bfloat16)
cfg = load_config("pickscore_cotrain_sd3_fast")
cfg.sample.num_batches_per_epoch = 1
cfg.train.gradient_accumulation_steps = 1
cfg.train_d = False
tr = GRPOTrainer(cfg, pipe, [f"synthetic prompt {i}" for i in range(9)], scorer=scorer, device=dev)
for _ in range(int(os.environ.get("WARM", "1"))):
    tr.run_epoch()
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.run_epoch()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one group")
