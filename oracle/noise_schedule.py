"""ORACLE (test infrastructure). Noise schedule of the reference sampler, restated for CPU/fp32.

  * SNR-shifted, log-interpolated betas        src/training/scheduler.py:16-30,32-58,100-120
    (hard-wired to mode="interpolated", shift_scale=6.0 at src/models/mvd_unet.py:417-428)
  * diffusers==0.32.2 DDPMScheduler (un-vendored; SD2.1 scheduler config: scaled_linear betas 0.00085..0.012,
    1000 train steps, v_prediction, fixed_small variance, leading spacing, steps_offset 1, no clipping):
    set_timesteps / step, as called at src/models/pipeline.py:119-120,161. Restated from SURVEY.md Appendix A.1.

The SNR/beta part is pinned against the reference's own scheduler.py (importable here) by oracle/gen_golden.py;
the DDPM step is PARITY UNPINNED (diffusers absent), checked only through its algebraic identities.
"""
from __future__ import annotations

from typing import Optional

import torch


def sd21_betas(n: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    """diffusers "scaled_linear": linspace(sqrt(b0), sqrt(b1), n) ** 2 in fp32."""
    return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2


def snr_from_betas(betas: torch.Tensor) -> torch.Tensor:
    """scheduler.py:32-58 with timesteps = arange(n): (sqrt(abar) / sqrt(1 - abar))^2."""
    abar = torch.cumprod(1.0 - betas, dim=0)
    return ((abar ** 0.5).float() / ((1.0 - abar) ** 0.5).float()) ** 2


def betas_from_snr(snr: torch.Tensor) -> torch.Tensor:
    """scheduler.py:16-30."""
    abar = ((snr / (1 + snr)) ** 0.5) ** 2
    alphas = abar / torch.cat([torch.ones(1), abar[:-1]])
    return 1 - alphas


def shifted_betas(betas: torch.Tensor, shift_scale: float = 6.0, mode: str = "interpolated") -> torch.Tensor:
    """scheduler.py:83-120: snr' = snr / s ("default") or exp((1-w) ln snr + w ln(snr/s)), w = t/(T-1)."""
    snr = snr_from_betas(betas)
    if mode == "default":
        return betas_from_snr(snr / shift_scale)
    n = betas.shape[0]
    w = torch.arange(n).float() / (n - 1)
    return betas_from_snr(torch.exp(torch.log(snr) * (1 - w) + torch.log(snr / shift_scale) * w))


class DDPMOracle:
    """DDPMScheduler(trained_betas=shifted, prediction_type="v_prediction", variance_type="fixed_small",
    timestep_spacing="leading", steps_offset=1, clip_sample=False)."""

    def __init__(self, betas: Optional[torch.Tensor] = None, steps_offset: int = 1):
        # from_config(..., trained_betas=numpy fp32) -> torch.tensor(trained_betas, dtype=float32)
        self.betas = (shifted_betas(sd21_betas()) if betas is None else betas).float()
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.num_train = self.betas.shape[0]
        self.steps_offset = steps_offset
        self.timesteps = torch.arange(self.num_train - 1, -1, -1)
        self.num_inference_steps = None

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = self.num_train // n
        self.timesteps = (torch.arange(0, n) * ratio).round().flip(0).long() + self.steps_offset
        return self.timesteps

    def coefficients(self, t: int):
        """(sqrt_abar, sqrt_1m_abar, c_x0, c_xt, sigma) for timestep t; sigma = 0 at t == 0."""
        ratio = self.num_train // self.num_inference_steps if self.num_inference_steps else 1
        prev = t - ratio
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else torch.tensor(1.0)
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        c_x0 = (a_prev ** 0.5 * cur_beta) / b_t
        c_xt = cur_alpha ** 0.5 * b_prev / b_t
        var = torch.clamp(b_prev / b_t * cur_beta, min=1e-20)
        sigma = var ** 0.5 if t > 0 else torch.tensor(0.0)
        return tuple(float(v) for v in (a_t ** 0.5, b_t ** 0.5, c_x0, c_xt, sigma))

    def step(self, model_output: torch.Tensor, t: int, sample: torch.Tensor, noise: Optional[torch.Tensor] = None):
        sa, sb, c0, ct, sg = self.coefficients(int(t))
        x0 = sa * sample - sb * model_output  # v-prediction
        prev = c0 * x0 + ct * sample
        if int(t) > 0 and noise is not None:
            prev = prev + sg * noise
        return prev


def denoise_loop(unet_fn, latents: torch.Tensor, sched: DDPMOracle, num_steps: int, guidance_scale: float = 1.0,
                 noises=None):
    """src/models/pipeline.py:140-166 — per step: CFG duplicate, UNet, CFG combine, scheduler.step.
    unet_fn(latent_model_input, t) -> model output; noises: per-step variance noise (injected, the reference
    calls scheduler.step without a generator)."""
    for i, t in enumerate(sched.set_timesteps(num_steps).tolist()):
        inp = torch.cat([latents] * 2) if guidance_scale > 1.0 else latents
        out = unet_fn(inp, t)
        if guidance_scale > 1.0:
            u, c = out.chunk(2)
            out = u + guidance_scale * (c - u)
        latents = sched.step(out, t, latents, None if noises is None else noises[i])
    return latents
