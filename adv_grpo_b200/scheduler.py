"""FlowMatchEulerDiscreteScheduler surface the reference touches (`pipeline.scheduler.{sigmas,
timesteps,index_for_timestep,set_timesteps,order,config}`; fast.py:574-580, sde.py:106-110).
SD3(.5) config: shift 3.0, 1000 train timesteps, no dynamic shifting.  The sigma/timestep tables
live on the device so the SDE-step kernel indexes them without any host synchronisation."""
import numpy as np
import torch


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, shift=3.0):
        self.config = type("Config", (), dict(num_train_timesteps=num_train_timesteps, shift=shift))()
        self.num_train_timesteps = num_train_timesteps
        self.shift = shift
        s = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        s = torch.from_numpy(s) / num_train_timesteps
        s = shift * s / (1 + (shift - 1) * s)
        self.sigma_min, self.sigma_max = s[-1].item(), s[0].item()
        self.timesteps = s * num_train_timesteps
        self.sigmas = torch.cat([s, torch.zeros(1)])
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, **_):
        N = self.num_train_timesteps
        if sigmas is None:
            ts = np.linspace(self.sigma_max * N, self.sigma_min * N, num_inference_steps)
            sigmas = ts / N
        sigmas = np.asarray(sigmas, dtype=np.float64)
        sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        sig = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        self.timesteps = sig * N
        self.sigmas = torch.cat([sig, torch.zeros(1, device=sig.device)])
        self.num_inference_steps = len(sig)

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        st = self.timesteps if schedule_timesteps is None else schedule_timesteps
        idx = (st == timestep).nonzero()
        return idx[1 if len(idx) > 1 else 0].item()


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, sigmas=None, **kw):
    scheduler.set_timesteps(num_inference_steps, device=device, sigmas=sigmas, **kw)
    return scheduler.timesteps, scheduler.num_inference_steps
