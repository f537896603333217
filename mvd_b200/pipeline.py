"""Drop-in for the denoising loop of the reference's `src/models/pipeline.py` (MVDPipeline.__call__ :11-186).

In scope (the hot path): the per-step body :140-166 — CFG duplication, `self.unet(...)`, CFG combine and
`scheduler.step` — executed as: torch.cat of the 4-channel latents (plumbing), the kernel-backed
MultiViewUNet, and ONE fused CFG+DDPM kernel updating the fp32 latents in place.
Out of scope: tokenizer + CLIP text encoder and the VAE. Callers pass
`prompt_embeds` (and `negative_prompt_embeds` for CFG) and `source_image_latents`; the result is the final
latents (`output_type="latent"`). Asking for prompts / images / PIL output raises.

Reference behaviours kept: the unconditional embedding is prepended only when guidance_scale > 1 AND one exists
(:64-83), while the latents are duplicated and the two halves combined whenever guidance_scale > 1 (:141,156-158;
without an unconditional embedding MultiViewUNet repeats the text over both halves); `scheduler.step` is ancestral DDPM (the reference passes no generator, :161) — here the per-step
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
                 variance_noises: Optional[torch.Tensor] = None, use_cuda_graph: bool = False):
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
        do_cfg = guidance_scale > 1.0  # reference :141,156: duplication and combine depend on the scale alone
        has_uncond = do_cfg and negative_prompt_embeds is not None  # reference :64-83
        if has_uncond:
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

        if use_cuda_graph:  # whole loop on the device: one captured step replayed num_inference_steps times
            sess = DenoiseSession(self, prompt_embeds[batch_size:] if has_uncond else prompt_embeds,
                                  num_inference_steps, guidance_scale,
                                  prompt_embeds[:batch_size] if has_uncond else None, source_camera, target_camera,
                                  source_image_latents, latent_size=tuple(latents.shape[-2:]))
            if variance_noises is None:
                variance_noises = torch.randn((sess.n_steps,) + tuple(latents.shape), device=dev, dtype=torch.float32,
                                              generator=generator)
            sess.reset(latents, variance_noises)
            latents = sess.run().clone()
            self.gpu_launch_count += sess.launches_per_step * sess.n_steps
            sess.close()
            return latents if not return_dict else {"images": latents, "latents": latents}

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


class DenoiseSession:
    """Device-resident sampling state of ONE object (its V views): static buffers plus, optionally, one CUDA graph
    holding a complete denoise step — CFG duplication, MultiViewUNet forward, fused CFG + DDPM update, step-counter
    advance — so the loop of reference pipeline.py:140-166 replays without host work.

    Per-step scalars (timestep, DDPM coefficients) and the per-step variance noise are read on the device from
    tables indexed by a device-side counter. The reference draws its positional projection on every UNet call;
    a session pins ONE draw for its lifetime (required for replay; pass `pos_proj` to choose it) and `close()`
    (also the context-manager exit) puts the encoder back to per-call draws if the session was the one that pinned it.
    `latent_size` is the latent height/width: an int (square) or an (H, W) pair.

    A captured step holds the addresses of every step-invariant tensor it read (weight packs, reference K/V, text
    K/V, camera embedding). Those caches are keyed on (data_ptr, _version), so a new source image, new cameras or
    reloaded weights must arrive as NEW tensors or through torch ops — and then `invalidate()` must be called so
    that the next step re-captures instead of replaying stale addresses."""

    def __init__(self, pipe: "MVDPipeline", prompt_embeds: torch.Tensor, num_inference_steps: int,
                 guidance_scale: float = 1.0, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 source_camera: Optional[torch.Tensor] = None, target_camera: Optional[torch.Tensor] = None,
                 source_image_latents: Optional[torch.Tensor] = None, latent_size=64,
                 use_cuda_graph: bool = True, pos_proj: Optional[torch.Tensor] = None, with_noise: bool = True):
        self.pipe, self.unet, dev = pipe, pipe.unet, pipe.device
        self.cfg = 2 if guidance_scale > 1.0 else 1  # reference pipeline.py:141,156
        self.guidance = float(guidance_scale)
        text = prompt_embeds.to(dev)
        if self.cfg == 2 and negative_prompt_embeds is not None:  # reference :79-81
            text = torch.cat([negative_prompt_embeds.to(dev), text])
        self.text = text.contiguous()
        self.views = prompt_embeds.shape[0]
        self.extra = {}
        if source_camera is not None:
            self.extra["source_camera"] = source_camera.to(dev).contiguous()
        if target_camera is not None:
            self.extra["target_camera"] = target_camera.to(dev).contiguous()
        if source_image_latents is not None:
            self.extra["source_image_latents"] = source_image_latents.to(dev).contiguous()
        cam = getattr(self.unet, "camera_encoder", None)
        self._unpin = None
        if cam is not None and target_camera is not None and (pos_proj is not None or cam._pos_proj is None):
            previous = cam._pos_proj
            cam.set_positional_projection(pos_proj if pos_proj is not None else cam._projection(dev).float())
            self._unpin = (cam, cam._pos_proj, previous)
        sched = pipe.scheduler
        sched.set_timesteps(num_inference_steps)
        self.timesteps = [int(t) for t in sched.timesteps]
        self.n_steps = len(self.timesteps)
        coef = torch.zeros(self.n_steps, 8)
        for i, t in enumerate(self.timesteps):
            coef[i, 0] = float(t)
            coef[i, 1:6] = torch.tensor(sched.coefficients(t))
        self.coef = coef.to(dev)
        lat_h, lat_w = (latent_size, latent_size) if isinstance(latent_size, int) else tuple(latent_size)
        self.latents = torch.zeros(self.views, 4, lat_h, lat_w, device=dev, dtype=torch.float32)
        self.noise_table = torch.zeros(self.n_steps, self.latents.numel(), device=dev, dtype=torch.float32) \
            if with_noise else None
        self.step_idx = torch.zeros(1, device=dev, dtype=torch.int32)
        self.t_dev = torch.full((1,), float(self.timesteps[0]), device=dev, dtype=torch.float32)
        # Everything that depends on the timestep only (Timesteps + TimestepEmbedding + the 22 time_emb_proj outputs)
        # is a per-SCHEDULE table, not per-step work: computed here once with the model's own kernels; the step reads
        # the current row, which advance_step_rows refreshes together with the step counter.
        base = getattr(self.unet, "base_unet", None)
        self.temb_table, self.temb_row = None, None
        if base is not None and hasattr(base, "timestep_rows") and next(base.parameters()).is_cuda:
            with torch.no_grad():
                rows = [base.timestep_rows(torch.full((1,), float(t), device=dev, dtype=torch.float32), 1)[1]
                        for t in self.timesteps]
            self.temb_table = torch.cat(rows, 0).contiguous()
            if self.temb_table.shape[1] % 4 == 0:
                self.temb_row = self.temb_table[0].clone()
            else:
                self.temb_table = None
        self.graph = None
        self.use_cuda_graph = use_cuda_graph
        self.launches_per_step = 0

    def invalidate(self):
        """Forget the captured step: the next `step()` warms the caches again and re-records. Call after anything the
        step treats as invariant was replaced (weights, source-image latents, cameras, text)."""
        self.graph = None

    def close(self):
        """Release the graph and, if this session pinned the camera encoder's positional projection, restore what was
        there before (None = the reference's per-call draws) — unless someone re-pinned it since."""
        self.graph = None
        if self._unpin is not None:
            cam, mine, previous = self._unpin
            if cam._pos_proj is mine:
                cam.set_positional_projection(previous)
            self._unpin = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def reset(self, latents: torch.Tensor, variance_noises: Optional[torch.Tensor] = None):
        """Load initial latents (and per-step noise [steps, V,4,L,L]); rewinds the step counter."""
        self.latents.copy_(latents.to(self.latents.device, torch.float32), non_blocking=True)
        if variance_noises is not None and self.noise_table is not None:
            self.noise_table.copy_(variance_noises.reshape(self.n_steps, -1).to(self.noise_table.device), non_blocking=True)
        self.step_idx.zero_()
        self.t_dev.fill_(float(self.timesteps[0]))
        if self.temb_row is not None:
            self.temb_row.copy_(self.temb_table[0])

    def forward_unet(self, inp: torch.Tensor) -> torch.Tensor:
        """The UNet call of the step, with the per-schedule time-embedding row installed for its duration."""
        base = getattr(self.unet, "base_unet", None)
        if self.temb_row is not None:
            base.temb_rows = self.temb_row
        try:
            return self.unet(sample=inp, timestep=self.t_dev, encoder_hidden_states=self.text, **self.extra).sample
        finally:
            if self.temb_row is not None:
                base.temb_rows = None

    def advance(self):
        if self.temb_row is not None:
            ops.advance_step_rows(self.step_idx, self.coef, self.t_dev, self.temb_table, self.temb_row)
        else:
            ops.advance_step(self.step_idx, self.coef, self.t_dev)

    def _eager_step(self):
        inp = torch.cat([self.latents] * 2) if self.cfg == 2 else self.latents
        out = self.forward_unet(inp)
        ops.cfg_ddpm_step_table(out, self.latents, self.noise_table, self.cfg, self.guidance, self.coef, self.step_idx)
        self.advance()

    # Local batches up to this many UNet samples launch their kernels with programmatic dependent launch. Mid-round it
    # paid for 1-2 samples per GPU (+5 % / +1 %, -3 % at 8); with stream-K GEMMs and persistent attention every kernel
    # already covers all SMs and the early CTAs of a dependent only get in the way: measured on the final tree
    # 4.565 (off) vs 4.652 ms at 1 sample, 5.80 vs 5.97 at 2, 8.02 vs 8.13 at 4, 12.52 vs 12.79 at 8 -> off (0).
    # MVD_PDL=1 still forces it on.
    LAUNCH_OVERLAP_MAX_BATCH = 0

    def capture(self, warmup: int = 2):
        """Warm every cache (weight packs, reference features, K/V) eagerly, then record one step."""
        ops.set_launch_overlap(self.views * self.cfg <= self.LAUNCH_OVERLAP_MAX_BATCH)
        saved = (self.latents.clone(), self.step_idx.clone(), self.t_dev.clone(),
                 None if self.temb_row is None else self.temb_row.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        before = ops.kernel_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # capture on the stream the warm-up ran on: the per-stream scratch buffers of ops.* (stream-K partials, attention
        # and GroupNorm workspaces) then already exist — created inside the capture, their zero-fill would become a
        # node of the graph and run on every replay (3 x 64 MiB fills = 26 us per step in the CUPTI timeline)
        with torch.cuda.graph(self.graph, stream=side):
            self._eager_step()
        self.launches_per_step = ops.kernel_launch_count() - before
        self.latents.copy_(saved[0])
        self.step_idx.copy_(saved[1])
        self.t_dev.copy_(saved[2])
        if saved[3] is not None:
            self.temb_row.copy_(saved[3])
        torch.cuda.synchronize()

    @torch.no_grad()
    def step(self):
        if self.use_cuda_graph:
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            before = ops.kernel_launch_count()
            self._eager_step()
            self.launches_per_step = ops.kernel_launch_count() - before

    @torch.no_grad()
    def run(self, steps: Optional[int] = None) -> torch.Tensor:
        for _ in range(self.n_steps if steps is None else steps):
            self.step()
        return self.latents

