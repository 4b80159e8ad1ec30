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


def _scaled_linear_alphas(num_train_timesteps, beta_start, beta_end, beta_schedule):
    if beta_schedule == "scaled_linear":
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                               dtype=torch.float32) ** 2
    elif beta_schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps,
                               dtype=torch.float32)
    else:
        raise NotImplementedError(beta_schedule)
    return torch.cumprod(1.0 - betas, dim=0)


class _SchedulerBase:
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, **kwargs):
        path = pretrained_model_name_or_path
        if subfolder:
            path = os.path.join(path, subfolder)
        with open(os.path.join(path, "scheduler_config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        cfg.update(kwargs)
        return cls(**cfg)


class DDPMScheduler(_SchedulerBase):
    """`add_noise` / `get_velocity` with per-(b,t,v) timesteps (reference
    src/dwm/schedulers/temporal_independent.py:8-45).  INT gather of the cumulative-alpha
    table, then one fused `s0[item]*x + s1[item]*y` pass on the GPU."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="linear", prediction_type="epsilon", **unused):
        self.config = _Config(num_train_timesteps=num_train_timesteps,
                              beta_start=beta_start, beta_end=beta_end,
                              beta_schedule=beta_schedule,
                              prediction_type=prediction_type)
        self.alphas_cumprod = _scaled_linear_alphas(
            num_train_timesteps, beta_start, beta_end, beta_schedule)

    def _mix(self, x, y, timesteps, sign):
        if not x.is_cuda:
            raise RuntimeError("scheduler kernels run on CUDA only (no CPU fallback)")
        self.alphas_cumprod = self.alphas_cumprod.to(x.device)
        a = self.alphas_cumprod[timesteps.to(x.device).long().flatten()]
        s0 = (a ** 0.5).float().contiguous()
        s1 = (sign * (1 - a) ** 0.5).float().contiguous()
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        _ops.lincomb2(x.float().contiguous(), y.float().contiguous(), s0, s1, out)
        return out.to(x.dtype)

    def add_noise(self, original_samples, noise, timesteps):
        # sqrt(a_t) * x0 + sqrt(1 - a_t) * noise
        return self._mix(original_samples, noise, timesteps, 1.0)

    def get_velocity(self, sample, noise, timesteps):
        # sqrt(a_t) * noise - sqrt(1 - a_t) * sample
        return self._mix(noise, sample, timesteps, -1.0)


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: torch.Tensor = None


class DDIMScheduler(_SchedulerBase):
    """DDIM (eta = 0) with tensor timesteps (reference
    src/dwm/schedulers/temporal_independent.py:48-170; `set_timesteps` and the alpha
    tables restate diffusers==0.31.0 DDIMScheduler)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="linear", clip_sample=True, set_alpha_to_one=True,
                 steps_offset=0, prediction_type="epsilon", thresholding=False,
                 timestep_spacing="leading", clip_sample_range=1.0, **unused):
        if thresholding or clip_sample:
            raise NotImplementedError(
                "clip_sample / thresholding are off in the SD-2.1 scheduler config")
        self.config = _Config(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start,
            beta_end=beta_end, beta_schedule=beta_schedule, clip_sample=clip_sample,
            set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
            prediction_type=prediction_type, timestep_spacing=timestep_spacing)
        self.alphas_cumprod = _scaled_linear_alphas(
            num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one \
            else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        n_train = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        spacing = self.config.timestep_spacing
        if spacing == "leading":
            ratio = n_train // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1]\
                .copy().astype(np.int64) + self.config.steps_offset
        elif spacing == "trailing":
            ratio = n_train / num_inference_steps
            ts = np.round(np.arange(n_train, 0, -ratio)).astype(np.int64) - 1
        elif spacing == "linspace":
            ts = np.linspace(0, n_train - 1, num_inference_steps).round()[::-1]\
                .copy().astype(np.int64)
        else:
            raise ValueError(spacing)
        self.timesteps = torch.from_numpy(ts).to(device)

    def step(self, model_output, timestep, sample, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None,
             variance_noise=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError(
                "Number of inference steps is 'None', you need to run "
                "'set_timesteps' after creating the scheduler")
        if eta != 0.0 or use_clipped_model_output:
            raise NotImplementedError("only the deterministic DDIM update (eta = 0)")
        if not model_output.is_cuda:
            raise RuntimeError("scheduler kernels run on CUDA only (no CPU fallback)")
        lead = sample.shape[:-3]
        ts = torch.as_tensor(timestep).to(sample.device, torch.int32)
        while ts.dim() > len(lead) and ts.shape[-1] == 1:
            ts = ts.squeeze(-1)
        ts = ts.expand(lead).contiguous()
        out = sample.to(torch.float32).contiguous().clone()
        _ops.cfg_ddim_step(
            model_output.float().contiguous(), out, ts,
            self.alphas_cumprod.to(sample.device), cfg=1, guidance_scale=1.0,
            step_ratio=self.config.num_train_timesteps // self.num_inference_steps,
            final_alpha_cumprod=float(self.final_alpha_cumprod),
            prediction_type=self.config.prediction_type)
        prev = out.to(sample.dtype)
        if not return_dict:
            return (prev, None)
        return DDIMSchedulerOutput(prev_sample=prev)
