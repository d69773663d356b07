"""Defaults read by the hot path (field names of the reference's `config/base.py:4-113`)."""
from . import ConfigDict


def get_config():
    c = ConfigDict()
    c.run_name = ""
    c.seed = 42
    c.logdir = "logs"
    c.save_freq = 20
    c.eval_freq = 20
    c.num_checkpoint_limit = 5
    c.mixed_precision = "bf16"
    c.allow_tf32 = True
    c.use_lora = True
    c.dataset = ""
    c.resolution = 512
    c.pretrained = ConfigDict(dict(model="stabilityai/stable-diffusion-3.5-medium", revision="main"))
    c.sample = ConfigDict(dict(num_steps=40, eval_num_steps=40, guidance_scale=4.5, train_batch_size=1,
                               num_image_per_prompt=1, test_batch_size=1, num_batches_per_epoch=2, global_std=True,
                               noise_level=0.7, same_latent=False))
    c.train = ConfigDict(dict(batch_size=1, use_8bit_adam=False, learning_rate=3e-4, adam_beta1=0.9, adam_beta2=0.999,
                              adam_weight_decay=1e-4, adam_epsilon=1e-8, gradient_accumulation_steps=1,
                              max_grad_norm=1.0, num_inner_epochs=1, cfg=True, adv_clip_max=5, clip_range=1e-4,
                              timestep_fraction=1.0, beta=0.0, lora_path=None, ema=False))
    c.prompt_fn = "general_ocr"
    c.prompt_fn_kwargs = {}
    c.reward_fn = ConfigDict()
    c.save_dir = ""
    c.per_prompt_stat_tracking = True
    return c
