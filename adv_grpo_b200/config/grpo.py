"""Presets of the two hot-path scripts (values of the reference's `config/grpo.py:315-376,389-430` and
its DINO twin), parameterised by the number of ranks instead of a hard-coded gpu_number=8."""
import os

from . import base


def _sd3_fast(gpu_number, reward_fn, dataset):
    c = base.get_config()
    c.dataset = os.path.join(os.getcwd(), dataset)
    c.mixed_precision = "bf16"
    c.wandb_init = False
    c.pretrained.model = "stabilityai/stable-diffusion-3.5-medium"
    c.sample.num_steps = 10
    c.sample.train_num_steps = 2
    c.sample.eval_num_steps = 40
    c.sample.guidance_scale = 4.5
    c.resolution = 512
    c.sample.train_batch_size = 1
    c.sample.num_image_per_prompt = 16
    c.sample.mini_num_image_per_prompt = 8
    c.sample.num_batches_per_epoch = int(48 / (gpu_number * c.sample.mini_num_image_per_prompt / c.sample.num_image_per_prompt))
    c.sample.test_batch_size = 16
    c.sample.random_timestep = 0
    c.train.batch_size = c.sample.mini_num_image_per_prompt
    c.train.gradient_accumulation_steps = max(c.sample.num_batches_per_epoch // 2, 1)
    c.train.num_inner_epochs = 1
    c.train.timestep_fraction = 0.99
    c.train.clip_range = 1e-5
    c.train.beta = 0.0
    c.sample.global_std = True
    c.sample.noise_level = 0.8
    c.train.ema = True
    c.save_freq = 60
    c.eval_freq = 60
    c.d_times = 20
    c.d_lr = 5e-6
    c.tune_layer = -1
    c.train_d = True
    c.weight_path = None
    c.json_path = None
    c.reference_image_path = None
    c.reward_fn = reward_fn
    c.eval_reward_fn = {"pickscore": 1}
    c.per_prompt_stat_tracking = True
    return c


def pickscore_cotrain_sd3_fast(gpu_number=8):
    c = _sd3_fast(gpu_number, {"pickscore_cotrain": 1}, "dataset/pickscore")
    c.discriminator = "pickscore"
    c.case_name = "fast_pickscore_cotrain"
    c.save_dir = "logs/pickscore/sd3.5-M-fast_pickscore_cotrain"
    return c


def dino_patch_cotrain_sd3_fast(gpu_number=8):
    c = _sd3_fast(gpu_number, {"dino_patch_cotrain": 1}, "dataset/pickscore")
    c.discriminator = "dino"
    c.d_times = 10
    c.d_lr = 1e-4
    c.case_name = "fast_dino_patch_cotrain"
    c.save_dir = "logs/dino/sd3.5-M-fast_dino_patch_cotrain"
    return c


def pickscore_sd3_fast(gpu_number=8):
    c = _sd3_fast(gpu_number, {"pickscore": 0.5, "ocr": 0.5}, "dataset/ocr")
    c.sample.random_timestep = None
    c.train_d = False
    c.case_name = "fast_multireward"
    c.save_dir = "logs/pickscore/sd3.5-M-fast_multireward_ocr_pickscore"
    return c


def get_config(name, **kw):
    return globals()[name](**kw)
