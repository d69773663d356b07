"""Oracle: FlowMatchEulerDiscreteScheduler (diffusers 0.33.1) as used by the
reference (`fast.py:574-580` retrieve_timesteps, `sde.py:106-110`
index_for_timestep / sigmas).  SD3(.5) scheduler config: shift=3.0,
num_train_timesteps=1000, use_dynamic_shifting=False.

Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np
import torch


class FlowMatchEulerOracle:
    def __init__(self, num_train_timesteps=1000, shift=3.0):
        self.num_train_timesteps = num_train_timesteps
        self.shift = shift
        ts = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        sig = torch.from_numpy(ts).to(torch.float32) / num_train_timesteps
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.sigma_min = sig[-1].item()
        self.sigma_max = sig[0].item()
        self.timesteps = sig * num_train_timesteps
        self.sigmas = sig
        self.order = 1

    def _sigma_to_t(self, sigma):
        return sigma * self.num_train_timesteps

    def set_timesteps(self, num_inference_steps, device=None):
        # diffusers: timesteps = linspace(sigma_to_t(sigma_max), sigma_to_t(sigma_min), n);
        # sigmas = timesteps / N; sigmas = shift*s/(1+(shift-1)*s)   (shift applied a 2nd time)
        """FlowMatchEulerDiscreteScheduler.set_timesteps via retrieve_timesteps (fast.py:574-580)."""
        ts = np.linspace(self._sigma_to_t(self.sigma_max), self._sigma_to_t(self.sigma_min),
                         num_inference_steps)
        sig = ts / self.num_train_timesteps
        sig = self.shift * sig / (1 + (self.shift - 1) * sig)
        sig = torch.from_numpy(sig).to(dtype=torch.float32, device=device)
        self.timesteps = sig * self.num_train_timesteps
        self.sigmas = torch.cat([sig, torch.zeros(1, device=sig.device)])
        self.num_inference_steps = num_inference_steps
        return self.timesteps

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        """FlowMatchEulerDiscreteScheduler.index_for_timestep as used at sd3_sde_with_logprob.py:106."""
        st = self.timesteps if schedule_timesteps is None else schedule_timesteps
        idx = (st == timestep).nonzero()
        pos = 1 if len(idx) > 1 else 0
        return idx[pos].item()
