"""ORACLE (test infrastructure, never shipped or timed as the product).

CPU/fp32 restatement of the arithmetic the reference delegates to the un-vendored third-party package
`diffusers==0.32.2` (pinned in the reference's uv.lock:722-723): `UNet2DConditionModel` with the public
`stabilityai/stable-diffusion-2-1` unet/config.json. The reference reaches this code at
src/models/mvd_unet.py:46-52,318-326 (main UNet) and src/models/image_encoder.py:18-22,105-110 (frozen
reference UNet). diffusers is not installable in this sandbox, so the published algorithm is restated from
its specification; module / parameter names are kept identical so that state-dict
keys match a real SD2.1 checkpoint and the reference's name walking
(src/models/mvd_unet.py:110-155, src/models/image_encoder.py:40-79) works unchanged.

PARITY UNPINNED for this file: the reference ships no tests/golden vectors and diffusers cannot be run
here; the structural known answers are the parameter count (865,910,724) and the FLOP totals of
SURVEY.md Appendix C, both asserted in tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import inspect
import math
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

SD21_CONFIG = dict(
    in_channels=4,
    out_channels=4,
    sample_size=96,
    block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2,
    cross_attention_dim=1024,
    attention_head_dim=(5, 10, 20, 20),  # diffusers naming quirk: these are HEAD COUNTS; dim_head = 64
    norm_num_groups=32,
    norm_eps=1e-5,
    down_has_attn=(True, True, True, False),
    up_has_attn=(False, True, True, True),
)


class _Config(dict):
    """attribute + item access, like diffusers' FrozenDict (src/models/mvd_unet.py:53,58,63 read attributes)."""

    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def timestep_sinusoid(timesteps: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half)
    ang = timesteps[:, None].float() * freqs[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    """GN -> SiLU -> conv3x3 -> +temb -> GN -> SiLU -> conv3x3 (+1x1 shortcut) -> residual."""

    def __init__(self, cin: int, cout: int, temb_dim: int = 1280, groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class AttnProcessor2_0:
    """diffusers AttnProcessor2_0: q/k/v projections -> SDPA -> out projection (no mask, no norm)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, *args,
                 **kwargs):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        b = hidden_states.shape[0]
        q = attn.to_q(hidden_states).view(b, -1, attn.heads, attn.dim_head).transpose(1, 2)
        k = attn.to_k(ctx).view(b, -1, attn.heads, attn.dim_head).transpose(1, 2)
        v = attn.to_v(ctx).view(b, -1, attn.heads, attn.dim_head).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, attn.heads * attn.dim_head).to(q.dtype)
        return attn.to_out[1](attn.to_out[0](o))


class Attention(nn.Module):
    def __init__(self, query_dim: int, heads: int, dim_head: int = 64, cross_attention_dim: Optional[int] = None):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        kv_dim = cross_attention_dim or query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        # diffusers Attention.forward: kwargs the processor's __call__ does not name are dropped.
        accepted = set(inspect.signature(self.processor.__call__).parameters.keys())
        kw = {k: v for k, v in cross_attention_kwargs.items() if k in accepted}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, cross_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, cross_attention_dim=cross_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, h, encoder_hidden_states=None, cross_attention_kwargs=None):
        kw = dict(cross_attention_kwargs or {})
        h = h + self.attn1(self.norm1(h), encoder_hidden_states=None, **kw)
        h = h + self.attn2(self.norm2(h), encoder_hidden_states=encoder_hidden_states, **kw)
        return h + self.ff(self.norm3(h))


class Transformer2DModel(nn.Module):
    def __init__(self, dim: int, heads: int, cross_dim: int, groups: int = 32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, encoder_hidden_states=None, cross_attention_kwargs=None, return_dict=False):
        b, c, hh, ww = x.shape
        h = self.norm(x).permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, encoder_hidden_states, cross_attention_kwargs)
        h = self.proj_out(h).reshape(b, hh, ww, c).permute(0, 3, 1, 2).contiguous()
        return (h + x,)


class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    """CrossAttnDownBlock2D (has_attn) / DownBlock2D."""

    def __init__(self, cin, cout, heads, cross_dim, has_attn, add_down, temb_dim=1280):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim) for i in range(2)])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim) for _ in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None
        self.has_attn = has_attn

    def forward(self, h, temb, encoder_hidden_states=None, cross_attention_kwargs=None):
        states = ()
        for i, res in enumerate(self.resnets):
            h = res(h, temb)
            if self.has_attn:
                h = self.attentions[i](h, encoder_hidden_states, cross_attention_kwargs)[0]
            states += (h,)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            states += (h,)
        return h, states


class MidBlock(nn.Module):
    """UNetMidBlock2DCrossAttn."""

    def __init__(self, c, heads, cross_dim, temb_dim=1280):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, cross_dim)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb_dim), ResnetBlock2D(c, c, temb_dim)])

    def forward(self, h, temb, encoder_hidden_states=None, cross_attention_kwargs=None):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, encoder_hidden_states, cross_attention_kwargs)[0]
        return self.resnets[1](h, temb)


class UpBlock(nn.Module):
    """CrossAttnUpBlock2D (has_attn) / UpBlock2D."""

    def __init__(self, cin, cout, prev_out, heads, cross_dim, has_attn, add_up, temb_dim=1280):
        super().__init__()
        res = []
        for i in range(3):
            skip = cin if i == 2 else cout
            inp = prev_out if i == 0 else cout
            res.append(ResnetBlock2D(inp + skip, cout, temb_dim))
        self.resnets = nn.ModuleList(res)
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim) for _ in range(3)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None
        self.has_attn = has_attn

    def forward(self, h, skips, temb, encoder_hidden_states=None, cross_attention_kwargs=None):
        for i, res in enumerate(self.resnets):
            h = torch.cat([h, skips[-1]], dim=1)
            skips = skips[:-1]
            h = res(h, temb)
            if self.has_attn:
                h = self.attentions[i](h, encoder_hidden_states, cross_attention_kwargs)[0]
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class UNetOut(tuple):
    """`.sample` + tuple behaviour, like diffusers' UNet2DConditionOutput / return_dict=False."""

    @property
    def sample(self):
        return self[0]


class UNet2DConditionModel(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(SD21_CONFIG)
        cfg.update(overrides)
        self.config = _Config(cfg)
        ch = cfg["block_out_channels"]
        heads = cfg["attention_head_dim"]
        cross = cfg["cross_attention_dim"]
        temb_dim = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg["in_channels"], ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb_dim)
        downs, out = [], ch[0]
        for i, c in enumerate(ch):
            downs.append(DownBlock(out, c, heads[i], cross, cfg["down_has_attn"][i], add_down=i < len(ch) - 1,
                                   temb_dim=temb_dim))
            out = c
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(ch[-1], heads[-1], cross, temb_dim)
        rev, rheads = list(reversed(ch)), list(reversed(heads))
        ups, out = [], rev[0]
        for i, c in enumerate(rev):
            prev_out, out = out, c
            cin = rev[min(i + 1, len(ch) - 1)]
            ups.append(UpBlock(cin, c, prev_out, rheads[i], cross, cfg["up_has_attn"][i], add_up=i < len(ch) - 1,
                               temb_dim=temb_dim))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(cfg["norm_num_groups"], ch[0], eps=cfg["norm_eps"])
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], cfg["out_channels"], 3, padding=1)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, sample, timestep, encoder_hidden_states, return_dict: bool = True, timestep_cond=None,
                cross_attention_kwargs: Optional[Dict[str, Any]] = None, added_cond_kwargs=None):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t.expand(sample.shape[0])
        temb = self.time_embedding(timestep_sinusoid(t, self.config.block_out_channels[0]).to(sample.dtype))
        h = self.conv_in(sample)
        skips: Tuple[torch.Tensor, ...] = (h,)
        for blk in self.down_blocks:
            h, st = blk(h, temb, encoder_hidden_states, cross_attention_kwargs)
            skips += st
        h = self.mid_block(h, temb, encoder_hidden_states, cross_attention_kwargs)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            h = blk(h, skips[-n:], temb, encoder_hidden_states, cross_attention_kwargs)
            skips = skips[:-n]
        h = self.conv_out(self.conv_act(self.conv_norm_out(h)))
        return UNetOut((h,))


def tiny_config() -> dict:
    """A structurally identical but small UNet (same block types, head_dim 64) for fast CPU parity tests."""
    return dict(block_out_channels=(64, 128, 128, 128), attention_head_dim=(1, 2, 2, 2), cross_attention_dim=64)
