"""DPM-Solver++ multistep scheduler — the `inference_config["scheduler"]` of the reference's
image example (`diffusers.DPMSolverMultistepScheduler`,
examples/ctsd_21_6views_image_generation.json; instantiated at
src/dwm/pipelines/ctsd.py:981-985, stepped with one scalar timestep per iteration at
:1573-1575).  Restates diffusers==0.31.0 for algorithm_type "dpmsolver++", solver_type
"midpoint", solver_order <= 2, no Karras sigmas / thresholding (diffusers' defaults on top of
the SD-2.1 `scheduler_config.json`).

All per-step coefficients depend only on the sigma table, so `set_timesteps` evaluates them
once (fp64 on the host) and keeps them on the device; `step` is then three launches of the
fused linear-combination kernel (`dwm_b200_lincomb2`) with no host-device traffic:

    x0   = c_x[i] * sample + c_m[i] * model_output          (epsilon / sample / v_prediction)
    prev = k_s[i] * sample + k_0[i] * x0_i + k_1[i] * x0_{i-1}
"""
import json
import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from opendwm_b200 import ops as _ops


class _Config(dict):
    __getattr__ = dict.get


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor


class DPMSolverMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="linear", trained_betas=None, solver_order=2,
                 prediction_type="epsilon", thresholding=False, algorithm_type="dpmsolver++",
                 solver_type="midpoint", lower_order_final=True, euler_at_final=False,
                 use_karras_sigmas=False, use_lu_lambdas=False, final_sigmas_type="zero",
                 lambda_min_clipped=-float("inf"), variance_type=None,
                 timestep_spacing="linspace", steps_offset=0, **unused):
        if algorithm_type != "dpmsolver++" or solver_type != "midpoint" or solver_order > 2 \
                or thresholding or use_karras_sigmas or use_lu_lambdas \
                or trained_betas is not None or lambda_min_clipped != -float("inf"):
            raise NotImplementedError(
                "only DPM-Solver++ (midpoint, order <= 2, plain sigmas) is provided")
        self.config = _Config(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, solver_order=solver_order,
            prediction_type=prediction_type, algorithm_type=algorithm_type,
            solver_type=solver_type, lower_order_final=lower_order_final,
            euler_at_final=euler_at_final, final_sigmas_type=final_sigmas_type,
            timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                   dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps,
                                   dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)
        self._begin_index = 0

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, **kwargs):
        path = pretrained_model_name_or_path
        if subfolder:
            path = os.path.join(path, subfolder)
        with open(os.path.join(path, "scheduler_config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        # a DDIM / PNDM style config (SD-2.1 ships one) only contributes the keys this class has
        keep = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule",
                "trained_betas", "prediction_type", "steps_offset", "timestep_spacing",
                "solver_order", "thresholding", "algorithm_type", "solver_type",
                "lower_order_final", "euler_at_final", "use_karras_sigmas",
                "use_lu_lambdas", "final_sigmas_type", "lambda_min_clipped")
        cfg = {k: v for k, v in cfg.items() if k in keep}
        if cfg.get("timestep_spacing") is None:
            cfg.pop("timestep_spacing", None)
        cfg.update(kwargs)
        return cls(**cfg)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index
        self._step_index = begin_index

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        last = c.num_train_timesteps
        n = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, last - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ratio = last // (n + 1)
            ts = (np.arange(0, n + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ratio = c.num_train_timesteps / n
            ts = np.arange(last, 0, -ratio).round().copy().astype(np.int64) - 1
        else:
            raise ValueError(c.timestep_spacing)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        last_sigma = 0.0 if c.final_sigmas_type == "zero" else \
            float(((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5)
        sigmas = np.concatenate([sig, [last_sigma]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)
        self.timesteps = torch.from_numpy(ts).to(device=device, dtype=torch.int64)
        self.num_inference_steps = n
        self.model_outputs = [None] * c.solver_order
        self.lower_order_nums = 0
        self._step_index = self._begin_index = 0
        self._coef = self._coefficients(sigmas.astype(np.float64)).to(device)

    def _coefficients(self, s):
        """[n, 2, 5] fp32: for step i and (first-order, second-order) variant the row
        (c_x, c_m, k_s, k_0, k_1) of the two linear combinations in the module docstring."""
        n = len(s) - 1
        out = np.zeros((n, 2, 5))
        alpha = 1.0 / np.sqrt(s * s + 1.0)
        sig = s * alpha

        def lam(j):
            return math.inf if sig[j] == 0 else math.log(alpha[j]) - math.log(sig[j])
        pt = self.config.prediction_type
        for i in range(n):
            if pt == "epsilon":
                cx, cm = 1.0 / alpha[i], -sig[i] / alpha[i]
            elif pt == "sample":
                cx, cm = 0.0, 1.0
            elif pt == "v_prediction":
                cx, cm = alpha[i], -sig[i]
            else:
                raise ValueError(pt)
            h = lam(i + 1) - lam(i)
            e1 = math.expm1(-h) if math.isfinite(h) else -1.0      # exp(-h) - 1
            ks, k = sig[i + 1] / sig[i], -alpha[i + 1] * e1
            out[i, 0] = (cx, cm, ks, k, 0.0)
            if i > 0:
                r0 = (lam(i) - lam(i - 1)) / h if math.isfinite(h) else 0.0
                out[i, 1] = (cx, cm, ks, k * (1.0 + 0.5 / r0), -0.5 * k / r0) if r0 != 0.0 \
                    else out[i, 0]
        return torch.from_numpy(out.astype(np.float32))

    def step(self, model_output, timestep=None, sample=None, generator=None,
             variance_noise=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run "
                             "'set_timesteps' after creating the scheduler")
        if not model_output.is_cuda:
            raise RuntimeError("scheduler kernels run on CUDA only (no CPU fallback)")
        i, c = self._step_index, self.config
        n = len(self.timesteps)
        lower_final = i == n - 1 and (c.euler_at_final or (c.lower_order_final and n < 15) or
                                      c.final_sigmas_type == "zero")
        first = c.solver_order == 1 or self.lower_order_nums < 1 or lower_final
        if self._coef.device != model_output.device:
            self._coef = self._coef.to(model_output.device)
        co = self._coef[i, 0 if first else 1]
        x = sample.to(torch.float32).contiguous()
        m = model_output.to(torch.float32).contiguous()
        x0 = torch.empty_like(x)
        _ops.lincomb2(x, m, co[0:1], co[1:2], x0)
        for k in range(c.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = x0
        prev = torch.empty_like(x)
        _ops.lincomb2(x, x0, co[2:3], co[3:4], prev)
        if not first:
            out = torch.empty_like(x)
            _ops.lincomb2(prev, self.model_outputs[-2], self._one(x.device), co[4:5], out)
            prev = out
        if self.lower_order_nums < c.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        prev = prev.to(model_output.dtype)
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev)

    def _one(self, device):
        one = self.__dict__.get("_one_t")
        if one is None or one.device != device:
            one = self._one_t = torch.ones(1, device=device)
        return one
