"""Drop-in for the reference's `src/models/mvd_unet.py`: `MultiViewUNet`, `UNetOutput`, `create_mvd_pipeline`
with the reference's constructor/forward keywords and attributes, orchestrating one denoise step
(reference mvd_unet.py:179-338) over the sm_100a kernels.

What is kept literally: the 32-processor installation and both name maps (mvd_unet.py:106-162), the FiLM forward
hooks on every down/mid/up block — including the reference's quirks that only `output[0]` of a down block is
modulated and that the mid hook is registered under the name "mid_0", which is not a modulator and therefore a
no-op (mvd_unet.py:357-376, camera_encoder.py:212-213) — the "output" modulator applied to the INPUT latents
(mvd_unet.py:256-258, fused into conv_in here), the text-embedding selection for the image encoder under CFG
(mvd_unet.py:278-285) and `UNetOutput`.

What is different by design: no `log_debug` tensor statistics (each cost a device->host sync in the reference,
src/utils.py:25-34), step-invariant work (reference-UNet features, reference/text K/V, camera embedding when the
positional projection is pinned) is cached, and `matched_batch_cfg=True` repeats per-view conditioning over the
CFG halves instead of the reference's flat re-view.
"""
from __future__ import annotations

from typing import Any, Dict, NamedTuple, Optional

import torch
import torch.nn as nn

from . import ops
from .attention import get_attention_processor_for_module
from .camera_encoder import CameraEncoder
from .image_encoder import ImageEncoder
from .scheduler import DDPMScheduler, ShiftSNRScheduler
import os

from .unet import BF16, UNet2DConditionModel, _small_linear_any_m, nhwc_view, nchw_shape

# apply the camera FiLM of the up blocks in the epilogue of the block's last GEMM / conv (MVD_FUSE_FILM=0: film kernel)
FUSE_UP_FILM = os.environ.get("MVD_FUSE_FILM", "1") != "0"


class UNetOutput(NamedTuple):
    sample: torch.Tensor


def _resolve_unet_config(pretrained_model_name_or_path) -> dict:
    """No network and no diffusers here: the UNet is built from the SD2.1 architecture (random init) and weights
    are loaded afterwards (MultiViewUNet.load_base_weights; keys match diffusers'). A dict overrides config entries."""
    if isinstance(pretrained_model_name_or_path, dict):
        return dict(pretrained_model_name_or_path)
    return {}


def _find_unet_weights(path) -> Optional[str]:
    """A local diffusers checkpoint directory (…/unet/diffusion_pytorch_model.{safetensors,bin}) or a weights file."""
    if not isinstance(path, (str, os.PathLike)):
        return None
    path = os.fspath(path)
    cands = [path] if os.path.isfile(path) else [
        os.path.join(path, sub, name) for sub in ("unet", "") for name in
        ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin")]
    for c in cands:
        if os.path.isfile(c):
            return c
    return None


def _read_state_dict(file: str) -> Dict[str, torch.Tensor]:
    if file.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(file)
    return torch.load(file, map_location="cpu", weights_only=True)


class MultiViewUNet(nn.Module):
    def __init__(self, pretrained_model_name_or_path=None, dtype: torch.dtype = torch.float32,
                 use_memory_efficient_attention: bool = True, enable_gradient_checkpointing: bool = True,
                 img_ref_scale: float = 0.3, cam_modulation_strength: float = 0.2, cam_output_dim: int = 1024,
                 cam_hidden_dim: int = 512, use_camera_conditioning: bool = True, use_image_conditioning: bool = True,
                 simple_cam_encoder: bool = False, matched_batch_cfg: bool = False,
                 cross_view_reference: bool = False):
        super().__init__()
        # cross-view mode (north star / BASELINE configs[3]): every sample attends over the reference tokens of ALL
        # views, passed to the processors as the 3-D reference [B, V*HW, C] of attention.py:95-132. Extension of the
        # reference's constructor; off by default.
        self.cross_view_reference = cross_view_reference
        self.use_camera_conditioning = use_camera_conditioning
        self.use_image_conditioning = use_image_conditioning
        self.cam_output_dim, self.cam_hidden_dim, self.simple_cam_encoder = cam_output_dim, cam_hidden_dim, simple_cam_encoder
        self.matched_batch_cfg = matched_batch_cfg
        cfg = _resolve_unet_config(pretrained_model_name_or_path)
        self.base_unet = UNet2DConditionModel(**cfg)
        self.config = self.base_unet.config
        self.device = torch.device("cpu")
        self.dtype = dtype
        self.img_ref_scale = img_ref_scale

        down_channels = list(self.config.block_out_channels)
        up_channels = list(reversed(down_channels))
        dims = {f"down_{i}": down_channels[min(i, len(down_channels) - 1)] for i in range(len(self.base_unet.down_blocks))}
        dims.update({f"up_{i}": up_channels[i] for i in range(len(self.base_unet.up_blocks))})
        dims["mid"] = down_channels[-1]
        dims["output"] = 4
        self.camera_encoder = CameraEncoder(output_dim=cam_output_dim, hidden_dim=cam_hidden_dim,
                                            modulation_hidden_dims=dims, modulation_strength=cam_modulation_strength,
                                            simple_encoder=simple_cam_encoder) if use_camera_conditioning else None
        self.image_encoder = ImageEncoder(pretrained_model_name_or_path, dtype=dtype,
                                          expected_sample_size=self.config.sample_size,
                                          unet_config=cfg) if use_image_conditioning else None
        self.hooks = []
        self.current_camera_embedding = None
        # multi-GPU view sharding (mvd_b200/dist.py): dict(view0, views_local, views_total[, cfg_total, ie_text])
        self.shard = None
        self.fuse_up_film = FUSE_UP_FILM  # camera FiLM of the up blocks applied in the block's last epilogue
        self._init_image_cross_attention()
        super().to(dtype=dtype)
        # reference mvd_unet.py:46-52 / image_encoder.py:18-22 load SD2.1 with from_pretrained and THEN seed the
        # adapters from those weights (attention.py:199-246). A local checkpoint is loaded the same way; a hub name
        # (no network here) or a missing path leaves RANDOM weights, which is said loudly instead of silently.
        if isinstance(pretrained_model_name_or_path, (str, os.PathLike)):
            found = _find_unet_weights(pretrained_model_name_or_path)
            if found is not None:
                self.load_base_weights(_read_state_dict(found))
            else:
                import warnings

                warnings.warn(f"MultiViewUNet: no local UNet weights under {pretrained_model_name_or_path!r} (this build "
                              "has no hub access): base UNet, image encoder and the adapters seeded from them are "
                              "RANDOM-INITIALISED. Call load_base_weights(sd21_unet_state_dict) before use.",
                              RuntimeWarning, stacklevel=2)

    def load_base_weights(self, state_dict: Dict[str, torch.Tensor], reseed_adapters: bool = True):
        """Load a diffusers SD2.1 `unet` state dict into the base UNet AND the frozen image-encoder UNet, then re-run
        the reference's adapter initialisation on all processors (attention.py:199-246: to_q_ref/to_k_ref/to_v_ref/
        to_out_ref copied from the — now pretrained — attention weights), exactly the order from_pretrained gives
        the reference. The processors' own keys (`*.processor.*`) are absent from an SD2.1 checkpoint, hence
        strict=False for them only; any other missing / unexpected key raises."""
        sd = {k: v for k, v in state_dict.items()}
        res = self.base_unet.load_state_dict(sd, strict=False)
        missing = [k for k in res.missing_keys if ".processor." not in k]
        if missing or res.unexpected_keys:
            raise RuntimeError(f"load_base_weights: missing {missing[:5]} unexpected {list(res.unexpected_keys)[:5]}")
        if self.image_encoder is not None:
            self.image_encoder.unet.load_state_dict(sd, strict=True)
        if reseed_adapters:
            with torch.no_grad():
                for attn in self.attention_layer_map.values():
                    attn.processor.load_original_weights(attn)
        return self

    _CACHE_KEYS = ("_pack_cache", "_ref_cache", "_gather_cache", "_ln_cache", "_ctx_cache", "_feat_cache",
                   "_text_cache", "_shard_idx", "_rep_cache", "_xview_cache", "_emb_cache", "_mod_cache", "_film_cache",
                   "_plist")

    def invalidate_caches(self):
        """Drop every step-invariant cache (weight packs, folded-LayerNorm packs, reference / text K/V, frozen-UNet
        features, camera embedding and FiLM coefficients) of this model, its image encoder, its camera encoder and all
        processors. The caches key on (data_ptr, _version) of their inputs, which the raw-pointer kernels of
        libmvd_b200 never bump: call this after refreshing such an input IN PLACE through ops.* (or an `out=` buffer),
        and `DenoiseSession.invalidate()` on any session whose captured step read the old tensors."""
        seen = 0
        owners = [m for m in self.modules()]
        owners += [getattr(a, "processor", None) for a in getattr(self, "attention_layer_map", {}).values()]
        for owner in owners:
            if owner is None:
                continue
            for key in self._CACHE_KEYS:
                if owner.__dict__.pop(key, None) is not None:
                    seen += 1
        return seen

    # ---- reference mvd_unet.py:106-162 ------------------------------------------------------------------------
    def _init_image_cross_attention(self):
        self.attention_layer_map = {}
        self.feature_to_attention_map = {}

        def install(prefix, attentions):
            for j, attn_block in enumerate(attentions):
                for tb in attn_block.transformer_blocks:
                    feature = f"{prefix}_attn_{j}"
                    names = []
                    for suffix, module in (("self", tb.attn1), ("cross", tb.attn2)):
                        name = f"{feature}_{suffix}"
                        self._replace_attention_processor(module, name)
                        names.append(name)
                    self.feature_to_attention_map[feature] = names

        for i, block in enumerate(self.base_unet.down_blocks):
            if hasattr(block, "attentions"):
                install(f"down_block_{i}", block.attentions)
        if hasattr(self.base_unet.mid_block, "attentions"):
            install("mid_block", self.base_unet.mid_block.attentions)
        for i, block in enumerate(self.base_unet.up_blocks):
            if hasattr(block, "attentions"):
                install(f"up_block_{i}", block.attentions)

    def _replace_attention_processor(self, attn_module, name):
        processor = get_attention_processor_for_module(name, attn_module, img_ref_scale=self.img_ref_scale)
        self.attention_layer_map[name] = attn_module
        attn_module.processor = processor

    def to(self, *args, **kwargs):
        device = args[0] if args and not isinstance(args[0], torch.dtype) else kwargs.get("device", self.device)
        self.device = torch.device(device) if device is not None else self.device
        if "dtype" in kwargs and kwargs["dtype"] is not None:
            self.dtype = kwargs["dtype"]
        if self.image_encoder is not None:
            self.image_encoder.device = self.device
        return super().to(*args, **kwargs)

    # ---- reference mvd_unet.py:179-338 ------------------------------------------------------------------------
    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor,
                source_camera: Optional[torch.Tensor] = None, target_camera: Optional[torch.Tensor] = None,
                source_image_latents: Optional[torch.Tensor] = None, return_dict: bool = True,
                timestep_cond=None, cross_attention_kwargs: Optional[Dict[str, Any]] = None, added_cond_kwargs=None):
        if cross_attention_kwargs and "debug_log_file_path" in cross_attention_kwargs:
            cross_attention_kwargs = {k: v for k, v in cross_attention_kwargs.items() if k != "debug_log_file_path"}
        dev = self.base_unet.device
        if dev.type != "cuda":
            raise RuntimeError("MultiViewUNet runs on a CUDA device only (no CPU path); call .to('cuda') first")
        sample = sample.to(device=dev)
        text = self._prepare_text(encoder_hidden_states.to(device=dev), sample.shape[0])

        self.current_camera_embedding = None
        self.base_unet.input_film = None
        if self.use_camera_conditioning and target_camera is not None:
            self._manage_modulation_hooks(register=True)
            self.current_camera_embedding = self.camera_encoder.encode_cameras(source_camera, target_camera)
            # "output" modulator on the input latents (mvd_unet.py:256-258): fused into conv_in
            mod = self.camera_encoder.modulation("output", self.current_camera_embedding)
            self.base_unet.input_film = (mod, float(self.camera_encoder.modulation_strength))
        # up blocks return only their output, so its FiLM is applied by the block's last kernel (conv / proj_out
        # epilogue) instead of a separate pass; down blocks also hand the UNMODULATED tensor to the skip stack
        # (camera_encoder.py:201-205), so theirs stays a kernel of its own
        for i, blk in enumerate(self.base_unet.up_blocks):
            blk.out_film = None
            if self.fuse_up_film and self.current_camera_embedding is not None:
                blk.out_film = self.camera_encoder.film_coefficients(f"up_{i}", self.current_camera_embedding,
                                                                     sample.shape[0])

        ref_hidden_states = None
        ref_batch_index = None
        if self.use_image_conditioning and source_image_latents is not None:
            batch_size = source_image_latents.shape[0]
            ie_text = text
            if text.shape[0] == 2 * batch_size:  # CFG: the conditional half (mvd_unet.py:280-283)
                ie_text = text[batch_size:]
            elif text.shape[0] > batch_size:
                ie_text = text[:batch_size]
            shard = self.shard
            sharded = False
            if shard is not None:
                # A rank owns views_local x cfg_local of the views_total x cfg_total samples of the object; it is
                # sharded as soon as EITHER dimension is split (one view with its CFG pair on two ranks included:
                # the unconditional rank must still feed the conditional text to the image encoder,
                # mvd_unet.py:280-283).
                cfg_total = int(shard.get("cfg_total", 2))
                if sample.shape[0] % shard["views_local"]:
                    raise ValueError(f"sharded step: local batch {sample.shape[0]} is not a multiple of views_local "
                                     f"{shard['views_local']}")
                cfg_local = sample.shape[0] // shard["views_local"]
                if cfg_local not in (1, cfg_total):
                    raise ValueError(f"sharded step: local batch holds {cfg_local} CFG branches of {cfg_total}")
                if batch_size != shard["views_total"]:
                    raise ValueError("sharded step: source_image_latents must hold ALL views of the object "
                                     f"({shard['views_total']}), got {batch_size}: the reference normalisation "
                                     "(attention.py:95-103) couples the batch")
                if cfg_total > 1 and not self.matched_batch_cfg:
                    raise ValueError("view/CFG sharding needs matched_batch_cfg=True: the reference-literal flat "
                                     "re-view of un-repeated features (attention.py:130-132) mixes the samples of "
                                     "different ranks")
                sharded = shard["views_local"] * cfg_local < shard["views_total"] * cfg_total
            xview_sharded = sharded and self.cross_view_reference
            if sharded and not xview_sharded:
                # features for ALL views (batch-coupled normalisation), conditional text of all views
                ie_text = self._prepare_text(shard["ie_text"].to(device=dev), batch_size)
            src_lat = source_image_latents.to(device=dev)
            emulated = bool(shard.get("emulated")) if shard is not None else False
            if xview_sharded:
                # cross-view mode: this rank encodes only ITS views; the token sets are all-gathered below
                # (emulated = a single-GPU stand-in for one rank: it encodes all views itself, no collective)
                ie_text = self._prepare_text(shard["ie_text"].to(device=dev), batch_size)
                if not emulated:
                    v0, vl = shard["view0"], shard["views_local"]
                    src_lat = src_lat[v0:v0 + vl]
                    ie_text = ie_text[v0:v0 + vl].contiguous()
            features = self.image_encoder(latents=src_lat, text_embeddings=ie_text, timestep=0)
            if xview_sharded:
                cfg_total = int(shard.get("cfg_total", 2))
                world = 1 if emulated else shard["views_total"] // shard["views_local"]
                features = self._cross_view_features(features, shard["views_total"] * cfg_total,
                                                     gather_group=shard.get("group"), world=world)
                ref_batch_index = None
            elif sharded:
                cfg_total = shard.get("cfg_total", 2)
                if self.matched_batch_cfg and cfg_total > 1:
                    features = self._repeat_features(features, cfg_total)
                ref_batch_index = self._shard_index(shard, sample.shape[0], dev)
            else:
                ref_batch_index = None
                if self.cross_view_reference:
                    features = self._cross_view_features(features, sample.shape[0])
                elif self.matched_batch_cfg and sample.shape[0] > batch_size:
                    features = self._repeat_features(features, sample.shape[0] // batch_size)
            ref_hidden_states = self._map_image_features_to_attention_layers(features)

        kw = dict(cross_attention_kwargs or {})
        if ref_hidden_states is not None:
            kw["ref_hidden_states"] = ref_hidden_states
            if ref_batch_index is not None:
                kw["ref_batch_index"] = ref_batch_index
        output = self.base_unet(sample=sample, timestep=timestep, encoder_hidden_states=text, return_dict=True,
                                timestep_cond=timestep_cond, cross_attention_kwargs=kw,
                                added_cond_kwargs=added_cond_kwargs)
        hidden_states = output.sample
        if not return_dict:
            return hidden_states
        return UNetOutput(sample=hidden_states)

    # ---- helpers ----------------------------------------------------------------------------------------------
    def _prepare_text(self, text: torch.Tensor, batch: int) -> torch.Tensor:
        """bf16 cast + CFG repeat (mvd_unet.py:233-237), cached per text tensor (constant across steps)."""
        key = (text.data_ptr(), text._version, tuple(text.shape), text.dtype, batch)
        cache = self.__dict__.setdefault("_text_cache", {})
        hit = cache.get(key)
        if hit is not None:
            return hit[0]
        t = text
        if t.dtype != BF16:
            t = ops.cast_bf16(t.float().contiguous())
        if batch > t.shape[0]:
            t = t.repeat(batch // t.shape[0], 1, 1)
        t = t.contiguous()
        if len(cache) >= 8:
            cache.clear()
        cache[key] = (t, text)  # keep `text` alive so its data_ptr cannot be recycled under the key
        return t

    def _shard_index(self, shard, local_batch: int, dev):
        from .dist import local_sample_index

        cfg_total = shard.get("cfg_total", 2)
        cfg_local = local_batch // shard["views_local"]
        key = (shard["view0"], shard["views_local"], cfg_local, shard.get("cfg_branch", 0))
        cached = self.__dict__.get("_shard_idx")
        if cached is None or cached[0] != key:
            idx = local_sample_index(shard["views_total"], cfg_total, shard["view0"], shard["views_local"], cfg_local,
                                     shard.get("cfg_branch", 0))
            cached = (key, torch.tensor(idx, device=dev, dtype=torch.long))
            self.__dict__["_shard_idx"] = cached
        return cached[1]

    def _repeat_features(self, features, times: int):
        key = (id(features), times)
        cached = self.__dict__.get("_rep_cache")
        if cached is not None and cached[0] == key and cached[2] is features:
            return cached[1]
        rep = {}
        for name, f in features.items():
            v = nhwc_view(f)
            rep[name] = nchw_shape(v.repeat(times, 1, 1, 1))
        self.__dict__["_rep_cache"] = (key, rep, features)
        return rep

    def _cross_view_features(self, features, batch: int, gather_group=None, world: int = 1):
        """[V, C, H, W] per site -> the 3-D reference of attention.py:95-132 holding the tokens of ALL views, the same
        for every one of the `batch` samples: ONE copy [1, V*HW, C] + its replication factor (SharedReference), never
        the [batch, V*HW, C] tensor itself. View-sharded ranks hold only their own views' features: those are
        all-gathered over NVLink (NCCL) — once per object, the features are step-invariant — so that every rank
        normalises and projects exactly the token set a single GPU would. Cached with the features."""
        from .attention import SharedReference

        key = (id(features), batch, world)
        cached = self.__dict__.get("_xview_cache")
        if cached is not None and cached[0] == key and cached[2] is features:
            return cached[1]
        out = {}
        for name, f in features.items():
            if isinstance(f, tuple):
                f = f[0]
            v = nhwc_view(f).contiguous()  # [V_local, H, W, C]
            if world > 1:
                import torch.distributed as dist

                full = torch.empty((world * v.shape[0],) + tuple(v.shape[1:]), device=v.device, dtype=v.dtype)
                dist.all_gather_into_tensor(full, v, group=gather_group)  # rank order == view order (shard_plan)
                v = full
            out[name] = SharedReference(v.reshape(1, -1, v.shape[-1]), batch)
        self.__dict__["_xview_cache"] = (key, out, features)
        return out

    def _map_image_features_to_attention_layers(self, image_features):
        ref_hidden_states = {}
        for name, feature in image_features.items():
            if isinstance(feature, tuple):
                feature = feature[0]
            for attn_name in self.feature_to_attention_map.get(name, []):
                if attn_name in self.attention_layer_map:
                    ref_hidden_states[attn_name] = feature
        return ref_hidden_states

    def _manage_modulation_hooks(self, register: bool):
        """reference mvd_unet.py:354-385 (hooks are registered once and stay)."""
        if register and not self.hooks:
            def get_hook(idx, direction):
                def hook(module, inputs, output):
                    if getattr(module, "film_applied", False):  # already applied in the block's last epilogue
                        module.film_applied = False
                        return output
                    if self.current_camera_embedding is not None:
                        return self.camera_encoder.apply_modulation(output, f"{direction}_{idx}",
                                                                    self.current_camera_embedding)
                    return output
                return hook

            targets = [(b, get_hook(i, "down")) for i, b in enumerate(self.base_unet.down_blocks)]
            targets.append((self.base_unet.mid_block, get_hook(0, "mid")))  # -> "mid_0": not a modulator, no-op
            targets += [(b, get_hook(i, "up")) for i, b in enumerate(self.base_unet.up_blocks)]
            for module, fn in targets:
                self.hooks.append(module.register_forward_hook(fn))
        elif not register and self.hooks:
            for h in self.hooks:
                h.remove()
            self.hooks = []


def create_mvd_pipeline(pretrained_model_name_or_path=None, dtype: torch.dtype = torch.bfloat16,
                        use_memory_efficient_attention: bool = True, enable_gradient_checkpointing: bool = True,
                        use_camera_conditioning: bool = True, use_image_conditioning: bool = True,
                        img_ref_scale: float = 0.25, cam_modulation_strength: float = 1.0, cam_output_dim: int = 1024,
                        cam_hidden_dim: int = 512, simple_cam_encoder: bool = False, cache_dir=None,
                        scheduler_config: Optional[Dict[str, Any]] = None, matched_batch_cfg: bool = False,
                        device: Optional[str] = None, cross_view_reference: bool = False):
    """reference mvd_unet.py:388-453: DDPM scheduler rebuilt on SNR-shifted (interpolated, scale 6) betas + the
    multi-view UNet. Text encoder and VAE are outside this library's scope (pass prompt_embeds / latents)."""
    from .pipeline import MVDPipeline

    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("mvd_b200 needs a CUDA device (sm_100a); there is no CPU path")
        device = "cuda"
    scheduler = ShiftSNRScheduler.from_scheduler(noise_scheduler=DDPMScheduler(), shift_mode="interpolated",
                                                 shift_scale=6.0, scheduler_class=DDPMScheduler)
    with torch.device(device):  # parameters are created (and initialised) directly on the GPU
        mv_unet = MultiViewUNet(pretrained_model_name_or_path, dtype=dtype,
                                use_memory_efficient_attention=use_memory_efficient_attention,
                                enable_gradient_checkpointing=enable_gradient_checkpointing,
                                img_ref_scale=img_ref_scale, cam_modulation_strength=cam_modulation_strength,
                                cam_output_dim=cam_output_dim, cam_hidden_dim=cam_hidden_dim,
                                simple_cam_encoder=simple_cam_encoder,
                                use_camera_conditioning=use_camera_conditioning,
                                use_image_conditioning=use_image_conditioning, matched_batch_cfg=matched_batch_cfg,
                                cross_view_reference=cross_view_reference)
    mv_unet = mv_unet.to(device=device, dtype=dtype)
    pipeline = MVDPipeline(unet=mv_unet, scheduler=scheduler)
    pipeline.use_camera_conditioning = use_camera_conditioning
    pipeline.use_image_conditioning = use_image_conditioning
    pipeline.img_ref_scale = img_ref_scale
    return pipeline
