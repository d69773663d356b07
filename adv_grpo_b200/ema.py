"""Exponential moving average of the trainable (LoRA) parameters -- the `EMAModuleWrapper` the training scripts
build at `train_sd3_fast_pickscore.py:528`, step once per sample batch (`:1186-1187`) and swap in for evaluation and
checkpoints (`:278-279,380-381,392-397`).  Interface of `adv_grpo/ema.py` (constructor, `step`, `copy_ema_to`,
`copy_temp_to`, `to`, `state_dict` / `load_state_dict`, attributes `ema_parameters` / `temp_stored_parameters`).

B200 layout: the generator has ONE flat fp32 LoRA parameter, so the shadow is one tensor and an update is one
multi-tensor lerp launch; nothing is staged through the host (the reference parks the swapped-out weights on the
CPU, `ema.py:66-67`; here they stay in HBM).

Schedule (`ema.py:33-37,45`): on steps where `(s + 1) % update_step_interval == 0`,
`shadow += (1 - min((1 + s) / (10 + s), decay)) * (param - shadow)` for parameters that require grad.
"""
import torch


def _detached_copies(tensors, device=None):
    return [t.detach().clone() if device is None else t.detach().clone().to(device) for t in tensors]


class EMAModuleWrapper:
    def __init__(self, parameters, decay=0.9999, update_step_interval=1, device=None):
        self.decay = decay
        self.update_step_interval = update_step_interval
        self.device = device
        self.ema_parameters = _detached_copies(list(parameters), device)
        self.temp_stored_parameters = None

    # ---- schedule ----------------------------------------------------------------------------------------
    def get_current_decay(self, optimization_step):
        warmup = (1 + optimization_step) / (10 + optimization_step)
        return warmup if warmup < self.decay else self.decay

    def _due(self, optimization_step):
        return (optimization_step + 1) % self.update_step_interval == 0

    # ---- update --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, parameters, optimization_step):
        live = list(parameters)
        if len(live) != len(self.ema_parameters):
            raise ValueError(f"EMA tracks {len(self.ema_parameters)} tensors, got {len(live)}")
        if not self._due(optimization_step):
            return
        shadow, current = [], []
        for s, p in zip(self.ema_parameters, live):
            if p.requires_grad:
                shadow.append(s)
                current.append(p.detach() if p.device == s.device else p.detach().to(s.device))
        if shadow:
            torch._foreach_lerp_(shadow, current, 1.0 - self.get_current_decay(optimization_step))

    # ---- swapping the averaged weights in and out ------------------------------------------------------------
    @torch.no_grad()
    def copy_ema_to(self, parameters, store_temp=True):
        live = list(parameters)
        if len(live) != len(self.ema_parameters):
            raise ValueError(f"EMA tracks {len(self.ema_parameters)} tensors, got {len(live)}")
        if store_temp:
            self.temp_stored_parameters = _detached_copies(live)
        for p, s in zip(live, self.ema_parameters):
            p.data.copy_(s.data, non_blocking=True)

    @torch.no_grad()
    def copy_temp_to(self, parameters):
        live = list(parameters)
        saved = self.temp_stored_parameters
        if saved is None or len(saved) != len(live):
            raise ValueError("copy_temp_to needs the weights stored by copy_ema_to(store_temp=True)")
        for p, t in zip(live, saved):
            p.data.copy_(t.data)
        self.temp_stored_parameters = None

    # ---- placement / persistence -------------------------------------------------------------------------------
    def to(self, device=None, dtype=None):
        self.device = device
        moved = []
        for s in self.ema_parameters:
            moved.append(s.to(device=device, dtype=dtype) if s.is_floating_point() else s.to(device=device))
        self.ema_parameters = moved

    def state_dict(self):
        return dict(decay=self.decay, ema_parameters=self.ema_parameters)

    def load_state_dict(self, state_dict):
        if not self.decay:
            self.decay = state_dict.get("decay", self.decay)
        self.ema_parameters = state_dict.get("ema_parameters")
        self.to(self.device)
