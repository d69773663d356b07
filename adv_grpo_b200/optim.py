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


class TorchOrderAdam:
    """`torch.optim.Adam(params, lr, betas)` of the discriminator steps (`scripts/train_sd3_fast_pickscore.py:658`,
    `scripts/train_sd3_fast_dino_patch.py:750`: lr = config.d_lr, betas = (0.5, 0.999), no weight decay) as one native
    pass per tensor (`csrc/heads.cu::adam_torch_order_kernel`).  The kernel follows torch's multi-tensor op order and
    rounds to the parameter dtype after every op, so a bf16 discriminator (the reference casts the DINO head and runs
    the PickScore model in bf16, with bf16 moments) takes the same steps as under torch -- including the updates that
    vanish below a bf16 ulp.  Surface: `param_groups`, `state`, `step()`, `zero_grad()`, `state_dict()` /
    `load_state_dict()` with torch's state names.  Parameters without a gradient are skipped, as torch does."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        params = [p for p in params]
        for p in params:
            if not p.is_cuda:
                raise ValueError("TorchOrderAdam drives CUDA parameters (there is no CPU fallback)")
        self.param_groups = [dict(params=params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=0.0)]
        self.state = {}

    def _state(self, p):
        st = self.state.get(p)
        if st is None:
            st = dict(step=0, exp_avg=torch.zeros_like(p, memory_format=torch.contiguous_format),
                      exp_avg_sq=torch.zeros_like(p, memory_format=torch.contiguous_format))
            self.state[p] = st
        return st

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        for p in g["params"]:
            if p.grad is None:
                continue
            if not p.is_contiguous():
                raise ValueError("TorchOrderAdam needs contiguous parameters")
            st = self._state(p)
            st["step"] += 1
            grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            ops.adam_torch_order(p.data, grad, st["exp_avg"], st["exp_avg_sq"], st["step"], g["lr"], g["betas"], g["eps"])
            torch.autograd.graph.increment_version(p)      # the kernel wrote through the raw pointer: caches keyed on
                                                           # p._version (packed weights, captured graphs) must see it

    def zero_grad(self, set_to_none=True):
        for p in self.param_groups[0]["params"]:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def state_dict(self):
        params = self.param_groups[0]["params"]
        state = {i: {"step": torch.tensor(float(self.state[p]["step"])), "exp_avg": self.state[p]["exp_avg"],
                     "exp_avg_sq": self.state[p]["exp_avg_sq"]} for i, p in enumerate(params) if p in self.state}
        groups = [{k: v for k, v in self.param_groups[0].items() if k != "params"} | {"params": list(range(len(params)))}]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        params = self.param_groups[0]["params"]
        for i, src in sd["state"].items():
            st = self._state(params[int(i)])
            st["step"] = int(src["step"])
            st["exp_avg"].copy_(src["exp_avg"])
            st["exp_avg_sq"].copy_(src["exp_avg_sq"])
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
