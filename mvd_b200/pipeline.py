"""Drop-in for the denoising loop of the reference's `src/models/pipeline.py` (MVDPipeline.__call__ :11-186).

In scope (the hot path): the per-step body :140-166 — CFG duplication, `self.unet(...)`, CFG combine and
`scheduler.step` — executed as: torch.cat of the 4-channel latents (plumbing), the kernel-backed
MultiViewUNet, and ONE fused CFG+DDPM kernel updating the fp32 latents in place.
Out of scope: tokenizer + CLIP text encoder and the VAE. Callers pass
`prompt_embeds` (and `negative_prompt_embeds` for CFG) and `source_image_latents`; the result is the final
latents (`output_type="latent"`). Asking for prompts / images / PIL output raises.

Reference behaviours kept: CFG is active only when guidance_scale > 1 AND an unconditional embedding exists
(:64-83); `scheduler.step` is ancestral DDPM (the reference passes no generator, :161) — here the per-step
variance noise comes from `generator` or from `variance_noises` (injected, for parity tests).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from . import ops


class MVDPipeline:
    def __init__(self, unet, scheduler, vae=None, text_encoder=None, tokenizer=None):
        self.unet, self.scheduler = unet, scheduler
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.vae_scale_factor = 8
        self.safety_checker = None
        self.feature_extractor = None
        self.gpu_launch_count = 0

    @property
    def device(self):
        return self.unet.base_unet.device

    def progress_bar(self, it):
        return it

    def prepare_latents(self, batch, channels, height, width, dtype, device, generator, latents=None):
        shape = (batch, channels, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=device, dtype=torch.float32)
        return latents * self.scheduler.init_noise_sigma

    @torch.no_grad()
    def __call__(self, prompt: Union[str, List[str], None] = None, height: Optional[int] = None,
                 width: Optional[int] = None, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                 negative_prompt=None, num_images_per_prompt: Optional[int] = 1, eta: float = 0.0,
                 generator: Optional[torch.Generator] = None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "latent", return_dict: bool = True,
                 callback: Optional[Callable[[int, int, torch.Tensor], None]] = None, callback_steps: int = 1,
                 cross_attention_kwargs: Optional[Dict[str, Any]] = None, source_camera: Optional[torch.Tensor] = None,
                 target_camera: Optional[torch.Tensor] = None, source_images: Optional[torch.Tensor] = None,
                 ref_scale: float = 0.1, use_camera_embeddings: bool = True, use_image_conditioning: bool = True,
                 debug_log_file_path: Optional[str] = None, source_image_latents: Optional[torch.Tensor] = None,
                 variance_noises: Optional[torch.Tensor] = None):
        if prompt_embeds is None:
            raise NotImplementedError("text encoding is outside mvd_b200's scope: pass prompt_embeds [B,77,1024]")
        if source_images is not None and source_image_latents is None:
            raise NotImplementedError("VAE encoding is outside mvd_b200's scope: pass source_image_latents")
        if output_type != "latent":
            raise NotImplementedError("VAE decoding is outside mvd_b200's scope: use output_type='latent'")
        if negative_prompt is not None and negative_prompt_embeds is None:
            raise NotImplementedError("pass negative_prompt_embeds instead of negative_prompt")
        dev = self.device
        batch_size = prompt_embeds.shape[0]
        do_cfg = guidance_scale > 1.0 and negative_prompt_embeds is not None  # reference :64-83
        if do_cfg:
            prompt_embeds = torch.cat([negative_prompt_embeds.to(dev), prompt_embeds.to(dev)])
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        latents = self.prepare_latents(batch_size * num_images_per_prompt, 4, height, width, torch.float32, dev,
                                       generator, None if latents is None else latents.to(dev, torch.float32))
        latents = latents.clone().contiguous()  # updated in place by the step kernel
        if source_image_latents is not None:
            source_image_latents = source_image_latents.to(dev)
            if source_image_latents.shape[0] < batch_size:  # reference :107-109
                source_image_latents = source_image_latents.repeat(batch_size // source_image_latents.shape[0], 1, 1, 1)

        self.scheduler.set_timesteps(num_inference_steps)
        timesteps = [int(t) for t in self.scheduler.timesteps]
        extra = {}
        if source_camera is not None:
            extra["source_camera"] = source_camera.to(dev)
        if target_camera is not None:
            extra["target_camera"] = target_camera.to(dev)
        if source_image_latents is not None:
            extra["source_image_latents"] = source_image_latents
        cross_attention_kwargs = cross_attention_kwargs or {}

        for i, t in enumerate(self.progress_bar(timesteps)):
            latent_model_input = torch.cat([latents] * 2) if do_cfg else latents
            noise_pred = self.unet(sample=latent_model_input, timestep=t, encoder_hidden_states=prompt_embeds,
                                   cross_attention_kwargs=cross_attention_kwargs, **extra).sample
            sa, sb, c0, ct, sg = self.scheduler.coefficients(t)
            noise = None
            if t > 0:
                noise = variance_noises[i].to(dev, torch.float32).contiguous() if variance_noises is not None else \
                    torch.randn(latents.shape, device=dev, dtype=torch.float32, generator=generator)
            # CFG combine (:156-158) + scheduler.step (:161) in one kernel, in place
            ops.cfg_ddpm_step(noise_pred, latents, noise, 2 if do_cfg else 1, float(guidance_scale), sa, sb, c0, ct, sg)
            if callback is not None and i % callback_steps == 0:
                callback(i, t, latents)
        if not return_dict:
            return latents
        return {"images": latents, "latents": latents}
