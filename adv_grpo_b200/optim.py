"""Optimizer of the generator update: global-norm clip + AdamW + gradient clear
as ONE native call (`csrc/optim.cu`) on the flat fp32 LoRA master parameter.

Replaces, with the same arithmetic, the three calls of `scripts/train_sd3_fast_pickscore.py:1165-1171`
(`accelerator.clip_grad_norm_(transformer.parameters(), max_grad_norm)`, `optimizer.step()`,
`optimizer.zero_grad()`; `optimizer = torch.optim.AdamW(lr, betas, weight_decay, eps)` at `:515-521`).
The class keeps the `torch.optim.Optimizer` surface the scripts touch: `param_groups` (so `lr` can be
scheduled), `step()`, `zero_grad()`, `state_dict()` / `load_state_dict()` with torch's AdamW state names.
"""
import torch

from . import ops


class FlatClipAdamW:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0):
        params = list(params)
        if len(params) != 1:
            raise ValueError("FlatClipAdamW drives one flat parameter (SD3Transformer2DModel.trainable_parameters())")
        p = params[0]
        if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
            raise ValueError("FlatClipAdamW needs a contiguous fp32 CUDA parameter (there is no CPU fallback)")
        self.param = p
        self.param_groups = [dict(params=[p], lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                  max_grad_norm=max_grad_norm)]
        self.state = {p: dict(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))}
        self.last_grad_norm = None

    @torch.no_grad()
    def step(self, zero_grad=True):
        """Clips (if `max_grad_norm > 0`), updates, and clears the gradient in the same pass unless `zero_grad=False`
        (then the gradient is left clipped, as `clip_grad_norm_` leaves it)."""
        p, g, st = self.param, self.param_groups[0], self.state[self.param]
        if p.grad is None:
            return None
        st["step"] += 1
        self.last_grad_norm = ops.clip_adamw(p.data, p.grad, st["exp_avg"], st["exp_avg_sq"], st["step"], g["lr"],
                                             g["betas"], g["eps"], g["weight_decay"], g["max_grad_norm"],
                                             zero_grad=zero_grad)
        return self.last_grad_norm

    def zero_grad(self, set_to_none=False):
        if self.param.grad is not None:
            if set_to_none:
                self.param.grad = None
            else:
                self.param.grad.zero_()

    def state_dict(self):
        st = self.state[self.param]
        groups = [{k: v for k, v in self.param_groups[0].items() if k != "params"} | {"params": [0]}]
        return {"state": {0: {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"],
                              "exp_avg_sq": st["exp_avg_sq"]}}, "param_groups": groups}

    def load_state_dict(self, sd):
        src, st = sd["state"][0], self.state[self.param]
        st["step"] = int(src["step"])
        st["exp_avg"].copy_(src["exp_avg"])
        st["exp_avg_sq"].copy_(src["exp_avg_sq"])
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
