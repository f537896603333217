"""Drop-in for the reference's `src/models/image_encoder.py`: the frozen second SD2.1 UNet run on the source-view
latents at t = 0, with forward hooks capturing the output of all 16 `Transformer2DModel`s
(reference image_encoder.py:36-84,97-112).

The reference re-runs this UNet on identical inputs at EVERY denoise step (mvd_unet.py:276,287-291: timestep
fixed at 0, latents and text constant). Here the feature dict is cached while the inputs are unchanged
 — bit-identical results, 804 GFLOP per view per step saved; `cache=False` restores the
reference's behaviour. Features are returned NCHW-shaped (channels-last memory), as the reference's are.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .unet import UNet2DConditionModel, _versions


def _key(t: torch.Tensor):
    return (t.data_ptr(), t._version, tuple(t.shape), t.dtype)


class ImageEncoder(nn.Module):
    def __init__(self, pretrained_model_name_or_path=None, dtype: torch.dtype = torch.float32,
                 expected_sample_size: int = None, unet_config: Optional[dict] = None, cache: bool = True):
        super().__init__()
        self.unet = UNet2DConditionModel(**(unet_config or {}))
        self.unet.config.sample_size = expected_sample_size
        for p in self.unet.parameters():
            p.requires_grad = False
        self.unet.eval()
        self.dtype = dtype
        self.device = "cpu"
        self.cache = cache
        self.extracted_features: Dict[str, torch.Tensor] = {}
        self._register_hooks()

    def _register_hooks(self):
        """reference image_encoder.py:36-79: one hook per attention block, named
        down_block_{i}_attn_{j} / mid_block_attn_{j} / up_block_{i}_attn_{j}."""
        self.hooks = []

        def add(name, layer):
            self.hooks.append(layer.register_forward_hook(lambda m, i, o, name=name: self._hook_fn(name, o)))

        for i, block in enumerate(self.unet.down_blocks):
            if hasattr(block, "attentions"):
                for j, layer in enumerate(block.attentions):
                    add(f"down_block_{i}_attn_{j}", layer)
        for j, layer in enumerate(self.unet.mid_block.attentions):
            add(f"mid_block_attn_{j}", layer)
        for i, block in enumerate(self.unet.up_blocks):
            if hasattr(block, "attentions"):
                for j, layer in enumerate(block.attentions):
                    add(f"up_block_{i}_attn_{j}", layer)

    def _hook_fn(self, name, output):
        self.extracted_features[name] = output[0] if isinstance(output, tuple) else output

    def to(self, *args, **kwargs):
        device = args[0] if args else kwargs.get("device", self.device)
        self.device = device
        if "dtype" in kwargs:
            self.dtype = kwargs["dtype"]
        return super().to(*args, **kwargs)

    def _unet_params(self):
        """The frozen UNet's parameters as a list built once (walking ~700 modules per eager call cost more than the
        cache lookup it serves); rebuilt when Module._apply (.to / .cuda / .half) may have replaced the tensors."""
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = list(self.unet.parameters())
            self.__dict__["_plist"] = plist
        return plist

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_plist", None)
        self.__dict__.pop("_feat_cache", None)
        return super()._apply(fn, *args, **kwargs)

    def forward(self, latents, text_embeddings, timestep):
        key = None
        if self.cache:
            t_key = _key(timestep) if torch.is_tensor(timestep) else float(timestep)
            key = (_key(latents), _key(text_embeddings), t_key, _versions(*self._unet_params()))
            cached = self.__dict__.get("_feat_cache")
            if cached is not None and cached[0] == key:
                self.extracted_features = cached[1]
                return self.extracted_features
        self.extracted_features = {}
        with torch.no_grad():
            self.unet(sample=latents, timestep=timestep, encoder_hidden_states=text_embeddings, return_dict=False)
        if self.cache:  # hold the inputs so their storage cannot be recycled under the same key
            self.__dict__["_feat_cache"] = (key, self.extracted_features, latents, text_embeddings, timestep)
        return self.extracted_features
