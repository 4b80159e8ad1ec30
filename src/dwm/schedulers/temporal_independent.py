"""Per-frame (temporally independent) schedulers — mirror of reference
src/dwm/schedulers/temporal_independent.py.

`FlowMatchEulerDiscreteScheduler` restates what the reference inherits from
diffusers==0.31.0 (`__init__`, `set_timesteps`: SURVEY.md Appendix A.8) and the
reference's own `step_by_indices` (:176-197).  The index arithmetic is integer and
bit-exact; the update itself runs in the fused CUDA kernel
(`dwm_b200_cfg_euler_step` in the pipeline, `dwm_b200_euler_step_by_indices` here).
"""
import json
import os
from dataclasses import dataclass

import numpy as np
import torch

from opendwm_b200 import ops as _ops


@dataclass
class FlowMatchEulerDiscreteSchedulerOutput:
    prev_sample: torch.Tensor


class _Config(dict):
    __getattr__ = dict.get


class FlowMatchEulerDiscreteScheduler:

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0,
                 use_dynamic_shifting: bool = False, **unused):
        if use_dynamic_shifting:
            raise NotImplementedError(
                "use_dynamic_shifting is not used by the SD-3.5 scheduler config")
        self.config = _Config(num_train_timesteps=num_train_timesteps,
                              shift=shift, use_dynamic_shifting=False)
        t = np.linspace(1, num_train_timesteps, num_train_timesteps,
                        dtype=np.float32)[::-1].copy()
        sigmas = torch.from_numpy(t).to(dtype=torch.float32) / num_train_timesteps
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = sigmas * num_train_timesteps
        self.sigmas = sigmas.to("cpu")
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()
        self.num_inference_steps = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None,
                        **kwargs):
        path = pretrained_model_name_or_path
        if subfolder:
            path = os.path.join(path, subfolder)
        with open(os.path.join(path, "scheduler_config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items()
                   if not k.startswith("_")}
        cfg.update(kwargs)
        return cls(**cfg)

    def _sigma_to_t(self, sigma):
        return sigma * self.config.num_train_timesteps

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        timesteps = np.linspace(
            self._sigma_to_t(self.sigma_max), self._sigma_to_t(self.sigma_min),
            num_inference_steps)
        sigmas = timesteps / self.config.num_train_timesteps
        shift = self.config.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        self.timesteps = (sigmas * self.config.num_train_timesteps)\
            .to(device=device)
        self.sigmas = torch.cat(
            [sigmas, torch.zeros(1, device=sigmas.device)])

    def step_by_indices(self, model_output: torch.FloatTensor, timestep_indices,
                        sample: torch.FloatTensor, return_dict: bool = True):
        """prev = sample + (sigma[idx+1] - sigma[idx]) * model_output in fp32, cast
        to model_output.dtype (reference :176-197).  timestep_indices: int tensor
        broadcastable over the leading dims of sample (e.g. [B, T, V])."""
        if not model_output.is_cuda:
            raise RuntimeError(
                "step_by_indices runs on CUDA only (no CPU fallback)")
        lead = sample.shape[:-3]
        idx = torch.as_tensor(timestep_indices).to(
            device=sample.device, dtype=torch.int32)
        while idx.dim() > len(lead) and idx.shape[-1] == 1:
            idx = idx.squeeze(-1)
        idx = idx.expand(lead).contiguous()
        out = sample.to(torch.float32).contiguous().clone()
        _ops.euler_step_by_indices(
            model_output.float().contiguous(), out, idx,
            self.sigmas.to(sample.device), round_dtype=model_output.dtype)
        prev_sample = out.to(model_output.dtype)
        if not return_dict:
            return (prev_sample,)
        return FlowMatchEulerDiscreteSchedulerOutput(prev_sample=prev_sample)


def df_timestep_indices(i: int, sequence_length: int, steps_per_inference: int,
                        take_time: int = 0):
    """Diffusion-forcing per-frame timestep index schedule (reference
    src/dwm/pipelines/ctsd.py:2048-2055).  Pure integer arithmetic."""
    return [
        min(i - take_time * steps_per_inference,
            max(0, i - j * steps_per_inference))
        for j in range(sequence_length)]


def df_in_schedule_range(i: int, sequence_length: int,
                         steps_per_inference: int):
    """reference src/dwm/pipelines/ctsd.py:2083-2088."""
    return [i - j * steps_per_inference >= 0 for j in range(sequence_length)]
