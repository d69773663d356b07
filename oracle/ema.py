"""Oracle: EMA of the trainable (LoRA) parameters (`adv_grpo/ema.py:39-49`).
Test infrastructure only (see oracle/__init__.py)."""


def ema_decay(decay, optimization_step):
    """EMAModuleWrapper.get_current_decay (ema.py:33-37)."""
    return min((1 + optimization_step) / (10 + optimization_step), decay)   # ema.py:33-37


def ema_step(ema_params, params, decay, update_step_interval, optimization_step):
    """EMAModuleWrapper.step (ema.py:39-56)."""
    omd = 1 - ema_decay(decay, optimization_step)
    if (optimization_step + 1) % update_step_interval == 0:                # ema.py:45
        for e, p in zip(ema_params, params):
            e.add_(omd * (p - e))                                           # ema.py:49
