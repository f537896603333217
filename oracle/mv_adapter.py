"""ORACLE (test infrastructure, never shipped or timed as the product).

CPU/fp32 restatement of MVD's adapter layers and of the per-step orchestration around the UNet, written from
the reference's behaviour (file:line cited per function; paths relative to the reference repository):

  * reference-image / cross-view attention branch      src/models/attention.py:12-188
  * processor construction and weight seeding           src/models/attention.py:199-265
  * camera encoder + FiLM                                src/models/camera_encoder.py:12-255
  * frozen reference UNet feature taps                   src/models/image_encoder.py:36-112
  * MultiViewUNet forward                                src/models/mvd_unet.py:63-80,106-162,229-385
  * denoise loop body (CFG + scheduler step)             src/models/pipeline.py:140-166

Pinning: the adapter parts (processor, camera encoder) ARE pinned — oracle/gen_golden.py imports the
reference's own attention.py / camera_encoder.py in the authoring container, checks this restatement
against them to fp32 round-off and stores the reference outputs under tests/golden/. The wrapper parts
that need diffusers (mvd_unet.py, image_encoder.py, pipeline.py) cannot be imported anywhere in this
sandbox: PARITY UNPINNED for those (structure only, see oracle/sd21_unet.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import math
from typing import Dict, List, NamedTuple, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .sd21_unet import UNet2DConditionModel


# --------------------------------------------------------------------------------------------------------
# reference-image cross attention
# --------------------------------------------------------------------------------------------------------
def normalize_reference(ref: torch.Tensor) -> torch.Tensor:
    """attention.py:95-103 — statistics over dims (0,1) of the RAW tensor (4-D: batch x channel per pixel,
    3-D: batch x sequence per channel), unbiased std clamped at 1e-6, times 0.5, under no_grad."""
    with torch.no_grad():
        r = ref.clone()
        r = r - r.mean(dim=(0, 1), keepdim=True)
        return r / torch.clamp(r.std(dim=(0, 1), keepdim=True), min=1e-6) * 0.5


def to_token_layout(ref: torch.Tensor) -> torch.Tensor:
    """attention.py:190-197 — NCHW -> [B, HW, C]; 3-D tensors pass through."""
    if ref.ndim == 4:
        b, c, h, w = ref.shape
        return ref.permute(0, 2, 3, 1).reshape(b, h * w, c)
    return ref


class RefAttnProcessor(nn.Module):
    """Same parameters (names, shapes) as the reference's ImageCrossAttentionProcessor (attention.py:33-43)."""

    def __init__(self, name: str, query_dim: int, heads: int, dim_head: int = 64, img_ref_scale: float = 0.3):
        super().__init__()
        inner = heads * dim_head
        self.name, self.heads, self.dim_head, self.query_dim = name, heads, dim_head, query_dim
        self.original_processor = None
        self.to_q_ref = nn.Linear(query_dim, inner, bias=False)
        self.to_k_ref = nn.Linear(query_dim, inner, bias=False)
        self.to_v_ref = nn.Linear(query_dim, inner, bias=False)
        self.ref_ln = nn.LayerNorm(inner)  # registered but unused (attention.py:37,160-161)
        self.to_out_ref = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])
        self.ref_scale_val = img_ref_scale

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 ref_hidden_states: Optional[Dict[str, torch.Tensor]] = None, *args, **kwargs):
        kwargs.pop("debug_log_file_path", None)
        base = self.original_processor(attn, hidden_states, encoder_hidden_states, attention_mask, temb=temb, *args,
                                       **kwargs)  # attention.py:62-70
        if ref_hidden_states is None or self.name not in ref_hidden_states:  # attention.py:72-81
            return base
        ref = to_token_layout(normalize_reference(ref_hidden_states[self.name]))
        b = hidden_states.shape[0]

        def split(t):  # attention.py:126,130,132 — flat view by the QUERY batch size
            return t.view(b, -1, self.heads, self.dim_head).transpose(1, 2)

        q = split(self.to_q_ref(hidden_states))
        k = split(self.to_k_ref(ref))
        v = split(self.to_v_ref(ref))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * self.dim_head)
        o = self.to_out_ref[1](self.to_out_ref[0](o))
        return base + self.ref_scale_val * o  # attention.py:174-181

    def seed_from(self, attn) -> None:
        """attention.py:199-246 — copy q/out; k/v copied when shapes agree, otherwise a TRANSPOSED slice of the
        text-attention weights (F.linear(eye(C), W[:C,:C]) == W[:C,:C]^T) or zero-padded columns."""
        with torch.no_grad():
            self.to_q_ref.weight.copy_(attn.to_q.weight)
            self.to_out_ref[0].weight.copy_(attn.to_out[0].weight)
            self.to_out_ref[0].bias.copy_(attn.to_out[0].bias)
            for mine, theirs in ((self.to_k_ref.weight, attn.to_k.weight), (self.to_v_ref.weight, attn.to_v.weight)):
                o, i = mine.shape
                oo, oi = theirs.shape
                if (o, i) == (oo, oi):
                    mine.copy_(theirs)
                elif i >= oi:
                    mine[:, :oi].copy_(theirs[: min(o, oo), :])
                    if i > oi:
                        mine[:, oi:].zero_()
                else:
                    mine.copy_(theirs[: min(o, oo), :i].t())


def make_processor(name: str, attn, img_ref_scale: float = 0.3) -> RefAttnProcessor:
    """attention.py:248-265."""
    heads = attn.heads
    proc = RefAttnProcessor(name, attn.to_q.in_features, heads, attn.to_q.out_features // heads, img_ref_scale)
    proc.original_processor = attn.processor
    proc.seed_from(attn)
    return proc


# --------------------------------------------------------------------------------------------------------
# camera encoder
# --------------------------------------------------------------------------------------------------------
def _mlp(dims: List[int]) -> nn.Sequential:
    layers: List[nn.Module] = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers += [nn.LayerNorm(dims[i + 1]), nn.SiLU()]
    return nn.Sequential(*layers)


class CameraEncoderOracle(nn.Module):
    """camera_encoder.py:12-105 (non-simple encoder). Same module/parameter names as the reference."""

    def __init__(self, output_dim=1024, hidden_dim=512, max_freq=10, modulation_hidden_dims=None,
                 modulation_strength=1.0):
        super().__init__()
        self.output_dim, self.max_freq = output_dim, max_freq
        self.pos_enc_dim = (output_dim // 2) // 3
        self.rotation_encoder = _mlp([9, hidden_dim, hidden_dim, output_dim])
        self.translation_encoder = _mlp([output_dim, hidden_dim, hidden_dim, output_dim])
        self.final_projection = nn.Sequential(nn.Linear(2 * output_dim, output_dim), nn.LayerNorm(output_dim), nn.SiLU(),
                                              nn.Linear(output_dim, output_dim), nn.LayerNorm(output_dim))
        self.output_norm = nn.LayerNorm(output_dim)
        self.modulators = nn.ModuleDict()
        for name, dim in (modulation_hidden_dims or {}).items():
            self.modulators[name] = _mlp([output_dim, output_dim // 2, dim * 2])
        for m in self.modulators.values():  # camera_encoder.py:92-105
            last = m[-1]
            nn.init.normal_(last.weight, mean=0.0, std=0.02)
            d = last.out_features // 2
            last.bias.data[:d].fill_(0.5)
            last.bias.data[d:].fill_(0.0)
        self.modulation_strength = modulation_strength

    @staticmethod
    def relative_transform(src: torch.Tensor, tgt: torch.Tensor):
        """camera_encoder.py:107-120."""
        rs, ts, rt, tt = src[:, :3, :3], src[:, :3, 3], tgt[:, :3, :3], tgt[:, :3, 3]
        r = torch.bmm(rt, rs.transpose(1, 2))
        return r, tt - torch.bmm(r, ts.unsqueeze(2)).squeeze(2)

    def sinusoid(self, t: torch.Tensor) -> torch.Tensor:
        """camera_encoder.py:137-151 (everything before the random projection)."""
        freqs = torch.exp(torch.linspace(0.0, math.log(self.max_freq), self.pos_enc_dim, device=t.device))
        ang = t.unsqueeze(-1) * freqs[None, None, :]
        return torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1).reshape(t.shape[0], -1)

    def encode_cameras(self, src, tgt, pos_proj: torch.Tensor) -> torch.Tensor:
        """camera_encoder.py:160-196. `pos_proj` [output_dim, 6*pos_enc_dim] stands for the matrix the reference
        draws with torch.randn(...)/sqrt(n) on EVERY call (camera_encoder.py:153-156); it is an explicit input here."""
        r, t = self.relative_transform(src.float(), tgt.float())
        rot = self.rotation_encoder(r.reshape(r.shape[0], -1))
        trans = self.translation_encoder(F.linear(self.sinusoid(t), pos_proj))
        return self.output_norm(self.final_projection(torch.cat([rot, trans], dim=-1)))

    def film(self, x: torch.Tensor, name: str, emb: torch.Tensor) -> torch.Tensor:
        """camera_encoder.py:211-234 on an NCHW tensor."""
        if name not in self.modulators:
            return x
        scale, shift = self.modulators[name](emb).chunk(2, dim=-1)
        scale = torch.sigmoid(scale)[:, :, None, None] * 2.0 * self.modulation_strength
        return x * scale + shift[:, :, None, None] * self.modulation_strength


# --------------------------------------------------------------------------------------------------------
# MultiViewUNet
# --------------------------------------------------------------------------------------------------------
class UNetOutput(NamedTuple):
    sample: torch.Tensor


def feature_taps(unet: UNet2DConditionModel):
    """image_encoder.py:36-79 — (name, Transformer2DModel) for every attention block, in registration order."""
    taps = []
    for i, blk in enumerate(unet.down_blocks):
        if hasattr(blk, "attentions"):
            taps += [(f"down_block_{i}_attn_{j}", m) for j, m in enumerate(blk.attentions)]
    taps += [(f"mid_block_attn_{j}", m) for j, m in enumerate(unet.mid_block.attentions)]
    for i, blk in enumerate(unet.up_blocks):
        if hasattr(blk, "attentions"):
            taps += [(f"up_block_{i}_attn_{j}", m) for j, m in enumerate(blk.attentions)]
    return taps


class MultiViewUNetOracle(nn.Module):
    """mvd_unet.py:22-385 with both SD2.1 UNets replaced by the restated one (random init unless loaded)."""

    def __init__(self, unet_config: Optional[dict] = None, img_ref_scale=0.3, cam_modulation_strength=0.2,
                 cam_output_dim=1024, cam_hidden_dim=512, use_camera_conditioning=True, use_image_conditioning=True,
                 matched_batch_cfg: bool = False, cross_view_reference: bool = False):
        super().__init__()
        cfg = unet_config or {}
        self.base_unet = UNet2DConditionModel(**cfg)
        self.config = self.base_unet.config
        self.use_camera_conditioning, self.use_image_conditioning = use_camera_conditioning, use_image_conditioning
        self.matched_batch_cfg = matched_batch_cfg  # repeat per-view conditioning over CFG halves
        # north-star / configs[3] mode: every sample attends over the reference tokens of ALL views, handed to the
        # processors as the 3-D reference [B, V*HW, C] that attention.py:95-132 accepts
        self.cross_view_reference = cross_view_reference
        ch = list(self.config.block_out_channels)
        dims = {f"down_{i}": ch[min(i, len(ch) - 1)] for i in range(len(self.base_unet.down_blocks))}
        dims.update({f"up_{i}": list(reversed(ch))[i] for i in range(len(self.base_unet.up_blocks))})
        dims["mid"] = ch[-1]
        dims["output"] = 4  # mvd_unet.py:63-80 — applied to the INPUT latents (mvd_unet.py:256-258)
        self.camera_encoder = CameraEncoderOracle(cam_output_dim, cam_hidden_dim, modulation_hidden_dims=dims,
                                                  modulation_strength=cam_modulation_strength) \
            if use_camera_conditioning else None
        self.image_encoder = None
        if use_image_conditioning:  # same attribute path as the reference: image_encoder.unet.*
            self.image_encoder = nn.Module()
            self.image_encoder.unet = UNet2DConditionModel(**cfg).requires_grad_(False).eval()
        # mvd_unet.py:106-162
        self.attention_layer_map, self.feature_to_attention_map = {}, {}
        for feat, tr in feature_taps(self.base_unet):
            for blk in tr.transformer_blocks:
                names = []
                for suffix, attn in (("self", blk.attn1), ("cross", blk.attn2)):
                    n = f"{feat}_{suffix}"
                    attn.processor = make_processor(n, attn, img_ref_scale)
                    self.attention_layer_map[n] = attn
                    names.append(n)
                self.feature_to_attention_map[feat] = names

    def image_features(self, latents, text) -> Dict[str, torch.Tensor]:
        """image_encoder.py:97-112 — frozen UNet at timestep 0; outputs of all 16 Transformer2DModels."""
        feats: Dict[str, torch.Tensor] = {}
        hooks = [m.register_forward_hook(lambda mod, i, o, n=n: feats.__setitem__(n, o[0] if isinstance(o, tuple) else o))
                 for n, m in feature_taps(self.image_encoder.unet)]
        try:
            with torch.no_grad():
                self.image_encoder.unet(latents, torch.tensor([0], dtype=torch.long), text, return_dict=False)
        finally:
            for h in hooks:
                h.remove()
        return feats

    def forward(self, sample, timestep, encoder_hidden_states, source_camera=None, target_camera=None,
                source_image_latents=None, pos_proj: Optional[torch.Tensor] = None, cross_attention_kwargs=None,
                return_features: bool = False):
        text = encoder_hidden_states
        if sample.shape[0] > text.shape[0]:  # mvd_unet.py:233-237
            text = text.repeat(sample.shape[0] // text.shape[0], 1, 1)
        emb = None
        if self.use_camera_conditioning and target_camera is not None:  # mvd_unet.py:241-258
            emb = self.camera_encoder.encode_cameras(source_camera, target_camera, pos_proj)
            if self.matched_batch_cfg and sample.shape[0] > emb.shape[0]:
                emb = emb.repeat(sample.shape[0] // emb.shape[0], 1)
            sample = self.camera_encoder.film(sample, "output", emb)
        ref = None
        if self.use_image_conditioning and source_image_latents is not None:  # mvd_unet.py:269-302
            nb = source_image_latents.shape[0]
            ie_text = text
            if text.shape[0] == 2 * nb:
                ie_text = text[nb:]
            elif text.shape[0] > nb:
                ie_text = text[:nb]
            feats = self.image_features(source_image_latents, ie_text)
            if self.cross_view_reference:
                feats = {k: v.flatten(2).transpose(1, 2).reshape(1, -1, v.shape[1]).repeat(sample.shape[0], 1, 1)
                         for k, v in feats.items()}
            elif self.matched_batch_cfg and sample.shape[0] > nb:
                feats = {k: v.repeat(sample.shape[0] // nb, 1, 1, 1) for k, v in feats.items()}
            ref = {a: f for n, f in feats.items() for a in self.feature_to_attention_map.get(n, [])}
        kw = dict(cross_attention_kwargs or {})
        if ref is not None:
            kw["ref_hidden_states"] = ref
        # mvd_unet.py:354-385 — FiLM forward hooks on every down / mid / up block output (first tuple element only)
        hooks = []
        if emb is not None:
            def mk(name):
                def hook(mod, inp, out):
                    if isinstance(out, tuple):
                        return (self.camera_encoder.film(out[0], name, emb),) + tuple(out[1:])
                    return self.camera_encoder.film(out, name, emb)
                return hook
            for i, b in enumerate(self.base_unet.down_blocks):
                hooks.append(b.register_forward_hook(mk(f"down_{i}")))
            hooks.append(self.base_unet.mid_block.register_forward_hook(mk("mid_0")))  # name "mid_0" ∉ modulators
            for i, b in enumerate(self.base_unet.up_blocks):
                hooks.append(b.register_forward_hook(mk(f"up_{i}")))
        try:
            out = self.base_unet(sample, timestep, text, cross_attention_kwargs=kw).sample
        finally:
            for h in hooks:
                h.remove()
        if return_features:
            return UNetOutput(sample=out), ref
        return UNetOutput(sample=out)
