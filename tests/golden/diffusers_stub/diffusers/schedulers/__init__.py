"""Base classes the reference's per-frame schedulers derive from
(/root/reference/src/dwm/schedulers/temporal_independent.py:6,48,173): the diffusers 0.31
alpha / sigma tables and `set_timesteps` restated in oracle/ (SURVEY.md Appendix A.8); the
`step`, `step_by_indices`, `add_noise`, `get_velocity` bodies that run are the REFERENCE's."""
from types import SimpleNamespace

import torch

from oracle import ctsd as _octsd
from oracle import d31 as _d31

from . import scheduling_ddim, scheduling_flow_match_euler_discrete  # noqa: F401


class _AlphaTables:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="linear", clip_sample=True, set_alpha_to_one=True,
                 steps_offset=0, prediction_type="epsilon", thresholding=False,
                 clip_sample_range=1.0, timestep_spacing="leading", **unused):
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, clip_sample=clip_sample,
            set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
            prediction_type=prediction_type, thresholding=thresholding,
            clip_sample_range=clip_sample_range, timestep_spacing=timestep_spacing)
        if beta_schedule == "scaled_linear":
            self.alphas_cumprod = _octsd.scaled_linear_alphas(
                num_train_timesteps, beta_start, beta_end)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps,
                                   dtype=torch.float32)
            self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        else:
            raise NotImplementedError(beta_schedule)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one \
            else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None


class DDPMScheduler(_AlphaTables):
    pass


class DDIMScheduler(_AlphaTables):

    def set_timesteps(self, num_inference_steps, device=None):
        assert self.config.timestep_spacing == "leading"
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        self.timesteps = ((torch.arange(0, num_inference_steps) * ratio).flip(0) +
                          self.config.steps_offset).to(device)

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 \
            else self.final_alpha_cumprod
        return (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)

    def _threshold_sample(self, sample):
        raise NotImplementedError("thresholding is off in the SD-2.1 scheduler config")


class FlowMatchEulerDiscreteScheduler(_d31.FlowMatchEulerDiscreteSchedulerBase):
    def __init__(self, num_train_timesteps=1000, shift=1.0, **unused):
        super().__init__(num_train_timesteps, shift)
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift)
        self._step_index = None

    def set_timesteps(self, num_inference_steps, device=None):
        super().set_timesteps(num_inference_steps, device)
        self._step_index = None

    def step(self, model_output, timestep, sample, return_dict=True, **unused):
        """diffusers 0.31 scalar-timestep Euler step (used by the reference's full-sequence
        loop, ctsd.py:1573-1575): the step index is found from the first timestep, then
        counted."""
        if self._step_index is None:
            self._step_index = int((self.timesteps == timestep).nonzero()[0].item())
        sample = sample.to(torch.float32)
        sigma, sigma_next = self.sigmas[self._step_index], self.sigmas[self._step_index + 1]
        prev = (sample + (sigma_next - sigma) * model_output).to(model_output.dtype)
        self._step_index += 1
        out = scheduling_flow_match_euler_discrete.FlowMatchEulerDiscreteSchedulerOutput(prev)
        return out if return_dict else (prev,)
