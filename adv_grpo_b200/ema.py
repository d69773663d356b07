"""Drop-in for `adv_grpo/ema.py` (`EMAModuleWrapper`): EMA of the trainable (LoRA) parameters with the
warm-up decay min((1+s)/(10+s), decay), updated every `update_step_interval` steps, swap-in/out for
eval and checkpointing.  Multi-tensor (`torch._foreach_*`) updates instead of a Python loop."""
import torch


class EMAModuleWrapper:
    def __init__(self, parameters, decay=0.9999, update_step_interval=1, device=None):
        parameters = list(parameters)
        self.ema_parameters = [p.clone().detach().to(device) for p in parameters]
        self.temp_stored_parameters = None
        self.decay, self.update_step_interval, self.device = decay, update_step_interval, device

    def get_current_decay(self, optimization_step):
        return min((1 + optimization_step) / (10 + optimization_step), self.decay)

    @torch.no_grad()
    def step(self, parameters, optimization_step):
        parameters = list(parameters)
        if len(parameters) != len(self.ema_parameters):
            raise ValueError("parameter list length changed")
        if (optimization_step + 1) % self.update_step_interval != 0:
            return
        w = 1 - self.get_current_decay(optimization_step)
        pairs = [(e, p) for e, p in zip(self.ema_parameters, parameters) if p.requires_grad]
        if not pairs:
            return
        ema = [e for e, _ in pairs]
        cur = [p.detach().to(e.device) for e, p in pairs]
        torch._foreach_lerp_(ema, cur, w)            # e += w (p - e)

    def to(self, device=None, dtype=None):
        self.device = device
        self.ema_parameters = [p.to(device=device, dtype=dtype) if p.is_floating_point() else p.to(device=device)
                               for p in self.ema_parameters]

    @torch.no_grad()
    def copy_ema_to(self, parameters, store_temp=True):
        parameters = list(parameters)
        if store_temp:
            self.temp_stored_parameters = [p.detach().clone() for p in parameters]
        for e, p in zip(self.ema_parameters, parameters, strict=True):
            p.data.copy_(e.to(p.device).data)

    @torch.no_grad()
    def copy_temp_to(self, parameters):
        for t, p in zip(self.temp_stored_parameters, parameters, strict=True):
            p.data.copy_(t.data)
        self.temp_stored_parameters = None

    def load_state_dict(self, state_dict):
        self.decay = self.decay if self.decay else state_dict.get("decay", self.decay)
        self.ema_parameters = state_dict.get("ema_parameters")
        self.to(self.device)

    def state_dict(self):
        return {"decay": self.decay, "ema_parameters": self.ema_parameters}
