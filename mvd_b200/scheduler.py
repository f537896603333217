"""Noise schedule of the reference sampler (host-side tables; the per-step arithmetic runs in
`mvd_cfg_ddpm_step_f32`).

  * ShiftSNRScheduler / compute_snr / SNR_to_betas      reference src/training/scheduler.py:16-58,74-150
  * DDPMScheduler (diffusers 0.32.2, un-vendored): SD2.1 scheduler config — scaled_linear betas
    0.00085..0.012, 1000 train steps, v_prediction, fixed_small variance, leading spacing, steps_offset 1 —
    as built at reference src/models/mvd_unet.py:417-428 and stepped at src/models/pipeline.py:119-120,161.
The tables are 1000 fp32 numbers computed once on the host with torch CPU ops (host logic, not the hot path).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Optional

import torch

from . import ops


def SNR_to_betas(snr: torch.Tensor) -> torch.Tensor:
    alpha_t = (snr / (1 + snr)) ** 0.5
    alphas_cumprod = alpha_t ** 2
    alphas = alphas_cumprod / torch.cat([torch.ones(1, device=snr.device), alphas_cumprod[:-1]])
    return 1 - alphas


def compute_snr(timesteps: torch.Tensor, noise_scheduler) -> torch.Tensor:
    acp = noise_scheduler.alphas_cumprod
    alpha = (acp ** 0.5)[timesteps].float()
    sigma = ((1.0 - acp) ** 0.5)[timesteps].float()
    return (alpha / sigma) ** 2


class DDPMScheduler:
    """The subset of diffusers' DDPMScheduler the reference pipeline uses."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", trained_betas=None, prediction_type: str = "v_prediction",
                 variance_type: str = "fixed_small", timestep_spacing: str = "leading", steps_offset: int = 1,
                 clip_sample: bool = False):
        if prediction_type != "v_prediction" or variance_type != "fixed_small" or timestep_spacing != "leading" \
                or clip_sample:
            raise NotImplementedError("only SD2.1's scheduler configuration is implemented")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, prediction_type=prediction_type,
                                      variance_type=variance_type, timestep_spacing=timestep_spacing,
                                      steps_offset=steps_offset, clip_sample=clip_sample)
        if trained_betas is not None:
            self.betas = torch.as_tensor(trained_betas, dtype=torch.float32).clone()
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    @classmethod
    def from_config(cls, config, **kwargs):
        base = dict(vars(config)) if not isinstance(config, dict) else dict(config)
        base.update(kwargs)
        return cls(**base)

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        self.timesteps = (torch.arange(0, num_inference_steps) * ratio).round().flip(0).long() + self.config.steps_offset
        if device is not None:
            self.timesteps = self.timesteps.to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, t: int):
        """(sqrt(abar_t), sqrt(1-abar_t), c_x0, c_xt, sigma_t) of DDPMScheduler.step at timestep t."""
        n = self.num_inference_steps or self.config.num_train_timesteps
        prev = t - self.config.num_train_timesteps // n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else torch.tensor(1.0)
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        sigma = torch.clamp(b_prev / b_t * cur_beta, min=1e-20) ** 0.5 if t > 0 else torch.tensor(0.0)
        return tuple(float(v) for v in (a_t ** 0.5, b_t ** 0.5, (a_prev ** 0.5 * cur_beta) / b_t,
                                        cur_alpha ** 0.5 * b_prev / b_t, sigma))

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, generator=None,
             variance_noise: Optional[torch.Tensor] = None, return_dict: bool = True):
        """diffusers signature. fp32 CUDA tensors; returns an object with `.prev_sample` (a new tensor)."""
        t = int(timestep)
        sa, sb, c0, ct, sg = self.coefficients(t)
        if t > 0 and variance_noise is None:
            variance_noise = torch.randn(sample.shape, device=sample.device, dtype=torch.float32, generator=generator)
        out = sample.float().clone()
        ops.cfg_ddpm_step(model_output.float().contiguous(), out, variance_noise if t > 0 else None, 1, 1.0, sa, sb,
                          c0, ct, sg)
        return SimpleNamespace(prev_sample=out) if return_dict else (out,)


class ShiftSNRScheduler:
    """reference src/training/scheduler.py:74-150."""

    def __init__(self, noise_scheduler: Any, timesteps: Any, shift_scale: float, scheduler_class: Any):
        self.noise_scheduler, self.timesteps = noise_scheduler, timesteps
        self.shift_scale, self.scheduler_class = shift_scale, scheduler_class

    def _rebuild(self, snr):
        return self.scheduler_class.from_config(self.noise_scheduler.config, trained_betas=SNR_to_betas(snr).numpy())

    def _get_shift_scheduler(self):
        return self._rebuild(compute_snr(self.timesteps, self.noise_scheduler) / self.shift_scale)

    def _get_interpolated_shift_scheduler(self):
        snr = compute_snr(self.timesteps, self.noise_scheduler)
        w = self.timesteps.float() / (self.noise_scheduler.config.num_train_timesteps - 1)
        return self._rebuild(torch.exp(torch.log(snr) * (1 - w) + torch.log(snr / self.shift_scale) * w))

    @classmethod
    def from_scheduler(cls, noise_scheduler: Any, shift_mode: str = "default", timesteps: Any = None,
                       shift_scale: float = 1.0, scheduler_class: Any = None):
        if timesteps is None:
            timesteps = torch.arange(0, noise_scheduler.config.num_train_timesteps)
        if scheduler_class is None:
            scheduler_class = noise_scheduler.__class__
        s = cls(noise_scheduler, timesteps, shift_scale, scheduler_class)
        if shift_mode == "default":
            return s._get_shift_scheduler()
        if shift_mode == "interpolated":
            return s._get_interpolated_shift_scheduler()
        raise ValueError(f"Unknown shift_mode: {shift_mode}")
