"""SD2.1 `UNet2DConditionModel` on sm_100a kernels.

The reference obtains this network from the un-vendored `diffusers==0.32.2` (reference
src/models/mvd_unet.py:46-52, src/models/image_encoder.py:18-22) and drives it at mvd_unet.py:318-326 /
image_encoder.py:105-110. This module keeps diffusers' module tree, attribute and parameter names
(state-dict keys of a real SD2.1 checkpoint load unchanged; the reference's name walking over
`down_blocks[i].attentions[j].transformer_blocks[k].attn1/attn2` works unchanged) and its calling
conventions (NCHW-shaped tensors between blocks, tuples from down blocks / Transformer2DModel, the attention
processor protocol with signature-filtered `cross_attention_kwargs`) — but every FLOP runs in
libmvd_b200.so:

  * activations are bf16 in channels-last memory (NCHW *shape*, NHWC *strides*), so the [B,HW,C] token view
    a transformer needs and the [B,H,W,C] view a convolution needs are the same bytes — no permute kernels;
  * conv3x3 / Linear -> tcgen05 implicit-GEMM (ops.conv3x3 / ops.linear) with bias, time-embedding,
    residual, GEGLU and skip-concatenation fused;
  * attention -> tcgen05 flash attention reading head slices in place from fused QKV projections.

There is no PyTorch compute fallback: CPU tensors or a missing shared library raise.
"""
from __future__ import annotations

import inspect
import os
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

BF16 = torch.bfloat16

SD21_CONFIG = dict(
    in_channels=4,
    out_channels=4,
    sample_size=96,
    block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2,
    cross_attention_dim=1024,
    attention_head_dim=(5, 10, 20, 20),  # head COUNTS (diffusers naming quirk); dim_head is 64 everywhere
    norm_num_groups=32,
    norm_eps=1e-5,
    down_has_attn=(True, True, True, False),
    up_has_attn=(False, True, True, True),
)

# tile width the GEGLU weight interleave is built for (ops.linear(..., geglu=True, tile_n=...)); 128 only for the
# experimental weight-stationary GEMM tiles (MVD_GEMM_WS=1), which are 128 wide
GEGLU_TILE = 128 if os.environ.get("MVD_GEGLU_TILE", "256") == "128" else 256
# fold the three LayerNorms of a transformer block into the GEMMs that consume them (MVD_FOLD_LN=0: LayerNorm kernels)
FOLD_LAYERNORM = os.environ.get("MVD_FOLD_LN", "1") != "0"
FOLD_GEGLU_MAX_ROWS = int(os.environ.get("MVD_FOLD_GEGLU_MAX_ROWS", "4096"))


class _Config(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


# ------------------------------------------------------------------------------------------------------------
# layout helpers: NCHW-shaped channels-last tensors <-> NHWC views
# ------------------------------------------------------------------------------------------------------------
def nhwc_view(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] channels-last -> contiguous [B,H,W,C] view (zero copy). Dense NCHW input is converted by the
    layout kernel (that only happens at the boundary, e.g. user-supplied reference features)."""
    if x.dim() != 4:
        raise ValueError("expected a 4-D tensor")
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    if not x.is_contiguous():
        raise ValueError("4-D activations must be dense NCHW or channels-last")
    b, c, h, w = x.shape
    if x.dtype not in (torch.float32, BF16):
        raise ValueError(f"unsupported dtype {x.dtype}")
    return ops.transpose_batched(x.view(b, c, h * w), out_dtype=BF16).view(b, h, w, c)


def nchw_shape(x_nhwc: torch.Tensor) -> torch.Tensor:
    """contiguous [B,H,W,C] -> [B,C,H,W]-shaped channels-last view (zero copy)."""
    return x_nhwc.permute(0, 3, 1, 2)


_SIG_CACHE: dict = {}


def accepted_params(processor) -> frozenset:
    """Parameter names of processor.__call__ (diffusers filters cross_attention_kwargs by them)."""
    key = type(processor)
    got = _SIG_CACHE.get(key)
    if got is None:
        got = frozenset(inspect.signature(processor.__call__).parameters.keys())
        _SIG_CACHE[key] = got
    return got


def _versions(*params) -> tuple:
    return tuple((p.data_ptr(), p._version) for p in params if p is not None)


class PackedModule(nn.Module):
    """nn.Module whose kernel-side weight layouts ("packs") are derived lazily from its parameters and rebuilt
    whenever a parameter is replaced or modified in place (load_state_dict, .to(), optimizer step)."""

    def _pack_params(self):
        raise NotImplementedError

    def _build_pack(self):
        raise NotImplementedError

    def pack(self):
        key = _versions(*self._pack_params())
        cached = self.__dict__.get("_pack_cache")
        if cached is None or cached[0] != key:
            with torch.no_grad():
                cached = (key, self._build_pack())
            self.__dict__["_pack_cache"] = cached
        return cached[1]


def _bf16(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t.detach().to(BF16).contiguous()


CONV_OUT_TENSOR_CORES = os.environ.get("MVD_CONV_OUT_TC", "1") != "0"


class LNInput:
    """A LayerNorm to be folded into the GEMM that consumes it: the module (gamma, beta, eps) + the row statistics of
    the raw activation, produced by the epilogue of the launch that wrote it (ops.RowStats)."""

    __slots__ = ("norm", "stats")

    def __init__(self, norm: nn.LayerNorm, stats):
        self.norm, self.stats = norm, stats


def fold_layernorm(w: torch.Tensor, norm: nn.LayerNorm, bias: Optional[torch.Tensor] = None):
    """LayerNorm(x) @ w^T + bias  ==  rstd * (x @ wg^T - mean * colsum) + c   with
    wg = w * gamma (per input channel, rounded to bf16 — what the MMA multiplies), colsum[n] = sum_k wg[n, k],
    c = w @ beta + bias. Returns (wg bf16 [N, K], colsum fp32 [N], c fp32 [1, N])."""
    wf = w.detach().float()
    wg = (wf * norm.weight.detach().float()[None, :]).to(BF16).contiguous()
    colsum = wg.float().sum(1).contiguous()
    c = wf @ norm.bias.detach().float()
    if bias is not None:
        c = c + bias.detach().float()
    return wg, colsum, c.view(1, -1).contiguous()


def conv3x3_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin] in (ky, kx, c) order (the K-major B operand of the implicit GEMM)."""
    co, ci = w.shape[:2]
    return w.detach().permute(0, 2, 3, 1).reshape(co, 9 * ci).to(BF16).contiguous()


# ------------------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------------------
class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, sin_emb: torch.Tensor) -> torch.Tensor:
        """sin_emb fp32 [B, in_dim] -> fp32 [B, dim] (skinny-linear kernel, SiLU fused)."""
        h = _small_linear_any_m(sin_emb, _bf16(self.linear_1.weight), _bf16(self.linear_1.bias), silu_out=True)
        return _small_linear_any_m(h, _bf16(self.linear_2.weight), _bf16(self.linear_2.bias))


def _small_linear_any_m(x, w, b, silu_in=False, silu_out=False):
    if x.shape[0] <= 16:
        return ops.small_linear(x, w, b, silu_in=silu_in, silu_out=silu_out)
    out = torch.empty((x.shape[0], w.shape[0]), device=x.device, dtype=torch.float32)
    for i in range(0, x.shape[0], 16):
        ops.small_linear(x[i:i + 16], w, b, silu_in=silu_in, silu_out=silu_out, out=out[i:i + 16])
    return out


class ResnetBlock2D(PackedModule):
    def __init__(self, cin: int, cout: int, temb_dim: int = 1280, groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.in_channels, self.out_channels, self.groups, self.eps = cin, cout, groups, eps
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def _pack_params(self):
        ps = [self.norm1.weight, self.norm1.bias, self.conv1.weight, self.conv1.bias, self.norm2.weight,
              self.norm2.bias, self.conv2.weight, self.conv2.bias, self.time_emb_proj.weight, self.time_emb_proj.bias]
        if self.conv_shortcut is not None:
            ps += [self.conv_shortcut.weight, self.conv_shortcut.bias]
        return ps

    def _build_pack(self):
        p = dict(
            g1=_bf16(self.norm1.weight), b1=_bf16(self.norm1.bias), w1=conv3x3_weight(self.conv1.weight),
            cb1=_bf16(self.conv1.bias), g2=_bf16(self.norm2.weight), b2=_bf16(self.norm2.bias),
            w2=conv3x3_weight(self.conv2.weight), cb2=_bf16(self.conv2.bias),
            tw=_bf16(self.time_emb_proj.weight), tb=_bf16(self.time_emb_proj.bias),
        )
        if self.conv_shortcut is not None:
            p["ws"] = _bf16(self.conv_shortcut.weight.reshape(self.out_channels, self.in_channels))
            p["bs"] = _bf16(self.conv_shortcut.bias)
        return p

    def forward(self, x: torch.Tensor, temb: torch.Tensor, skip: Optional[torch.Tensor] = None,
                temb_proj: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x (and optional `skip`, concatenated after x along C — diffusers does torch.cat before the block):
        NCHW-shaped channels-last bf16. temb: fp32 [B,1280]; temb_proj: precomputed time_emb_proj(SiLU(temb))."""
        p = self.pack()
        xa = nhwc_view(x)
        xb = nhwc_view(skip) if skip is not None else None
        if temb_proj is None:
            temb_proj = _small_linear_any_m(temb, p["tw"], p["tb"], silu_in=True)
        h = ops.groupnorm(xa, p["g1"], p["b1"], self.groups, self.eps, silu=True, x2=xb)
        h = ops.conv3x3(h, p["w1"], bias=p["cb1"], img_bias=temb_proj)
        h = ops.groupnorm(h, p["g2"], p["b2"], self.groups, self.eps, silu=True)
        n, hh, ww, _ = xa.shape
        if self.conv_shortcut is not None:
            a2 = xb.reshape(n * hh * ww, -1) if xb is not None else None
            res = ops.linear(xa.reshape(n * hh * ww, -1), p["ws"], bias=p["bs"], a2=a2).view(n, hh, ww, -1)
        else:
            if xb is not None:
                raise ValueError("a skip concatenation always changes the channel count; conv_shortcut expected")
            res = xa
        out = ops.conv3x3(h, p["w2"], bias=p["cb2"], residual=res)
        return nchw_shape(out)


class AttnProcessor2_0:
    """diffusers AttnProcessor2_0 (the reference's `original_processor`, src/models/attention.py:62-70,261):
    to_out(SDPA(to_q h, to_k e, to_v e)). Fused QKV projection, in-place head slicing, optional fused residual."""

    def __call__(self, attn: "Attention", hidden_states: torch.Tensor, encoder_hidden_states=None,
                 attention_mask=None, temb=None, residual: Optional[torch.Tensor] = None,
                 ln_fold: Optional[LNInput] = None, want_stats: bool = False, *args, **kwargs):
        """ln_fold: hidden_states is the RAW residual stream and the block's LayerNorm is folded into the input
        projection (LNInput); want_stats: also return the ops.RowStats of the result (for the next LayerNorm)."""
        if attention_mask is not None:
            raise NotImplementedError("attention masks are not on MVD's hot path")
        b, s, c = hidden_states.shape
        pk = attn.pack()
        hs2d = hidden_states.reshape(b * s, c)
        w_in = pk["wqkv"] if encoder_hidden_states is None else pk["wq"]
        if ln_fold is not None:
            wg, colsum, cst = attn.ln_pack(ln_fold.norm, "wqkv" if encoder_hidden_states is None else "wq")
            proj = ops.linear(hs2d, wg, row_group_bias=cst, rows_per_group=b * s,
                              ln=ops.LNFold(ln_fold.stats, colsum, ln_fold.norm.eps))
        else:
            proj = ops.linear(hs2d, w_in)
        if encoder_hidden_states is None:
            qkv = proj.view(b, s, 3 * c)
            q, k, v = qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:]
        else:
            q = proj.view(b, s, c)
            kv = attn.context_kv(encoder_hidden_states)
            k, v = kv[:, :, :c], kv[:, :, c:]
        o = ops.attention(q, k, v, attn.heads, attn.scale)
        res2d = residual.reshape(b * s, c) if residual is not None else None
        out = ops.linear(o.view(b * s, c), pk["wo"], bias=pk["bo"], residual=res2d, want_stats=want_stats)
        if want_stats:
            return out[0].view(b, s, c), out[1]
        return out.view(b, s, c)


class Attention(PackedModule):
    def __init__(self, query_dim: int, heads: int, dim_head: int = 64, cross_attention_dim: Optional[int] = None):
        super().__init__()
        if dim_head != 64:
            raise ValueError("the attention kernel is specialised for head_dim 64 (SD2.1)")
        inner = heads * dim_head
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        self.is_cross = cross_attention_dim is not None
        kv_dim = cross_attention_dim or query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def _pack_params(self):
        return [self.to_q.weight, self.to_k.weight, self.to_v.weight, self.to_out[0].weight, self.to_out[0].bias]

    def _build_pack(self):
        p = dict(wq=_bf16(self.to_q.weight), wo=_bf16(self.to_out[0].weight), bo=_bf16(self.to_out[0].bias),
                 wkv=_bf16(torch.cat([self.to_k.weight, self.to_v.weight], 0)))
        if not self.is_cross:
            p["wqkv"] = _bf16(torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], 0))
        self.__dict__["_ctx_cache"] = None
        return p

    def ln_pack(self, norm: nn.LayerNorm, which: str):
        """(gamma-scaled weight, column sums, W.beta) of the input projection `which` for a preceding LayerNorm."""
        pk = self.pack()
        key = (which, _versions(*self._pack_params(), norm.weight, norm.bias))
        cache = self.__dict__.setdefault("_ln_cache", {})
        hit = cache.get(which)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, fold_layernorm(pk[which], norm))
            cache[which] = hit
        return hit[1]

    def context_kv(self, ctx: torch.Tensor) -> torch.Tensor:
        """[B, S_ctx, 2C] = [to_k(ctx) | to_v(ctx)]; cached while the same context tensor is passed (text
        embeddings are constant across denoise steps)."""
        key = (ctx.data_ptr(), ctx._version, tuple(ctx.shape), ctx.dtype)
        pk = self.pack()
        cached = self.__dict__.get("_ctx_cache")
        if cached is not None and cached[0] == key:
            return cached[1]
        b, s, d = ctx.shape
        c2d = ctx.reshape(b * s, d)
        if c2d.dtype != BF16:
            c2d = ops.cast_bf16(c2d.float().contiguous())
        kv = ops.linear(c2d, pk["wkv"]).view(b, s, -1)
        self.__dict__["_ctx_cache"] = (key, kv, ctx)  # keep ctx alive so data_ptr cannot be recycled
        return kv

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        # diffusers Attention.forward: kwargs not named by the processor's __call__ are dropped
        accepted = accepted_params(self.processor)
        kw = {k: v for k, v in cross_attention_kwargs.items() if k in accepted}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)


class FeedForward(PackedModule):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim)])

    def _pack_params(self):
        return [self.net[0].proj.weight, self.net[0].proj.bias, self.net[2].weight, self.net[2].bias]

    def _build_pack(self):
        w, b = self.net[0].proj.weight.detach(), self.net[0].proj.bias.detach()
        inner = w.shape[0] // 2
        half = GEGLU_TILE // 2
        if inner % half:
            raise ValueError(f"GEGLU inner dim {inner} must be a multiple of {half}")
        # per 256-column tile: [128 value rows | 128 gate rows] so the epilogue pairs them inside one tile
        wp = torch.cat([w[:inner].reshape(-1, half, w.shape[1]), w[inner:].reshape(-1, half, w.shape[1])], 1)
        bp = torch.cat([b[:inner].reshape(-1, half), b[inner:].reshape(-1, half)], 1)
        return dict(w1=_bf16(wp.reshape(2 * inner, -1)), b1=_bf16(bp.reshape(-1)), w2=_bf16(self.net[2].weight),
                    b2=_bf16(self.net[2].bias))

    def ln_pack(self, norm: nn.LayerNorm):
        p = self.pack()
        key = _versions(*self._pack_params(), norm.weight, norm.bias)
        hit = self.__dict__.get("_ln_cache")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                wg, colsum, cst = fold_layernorm(p["w1"], norm, p["b1"])  # rows already in the GEGLU interleave
                hit = (key, (wg, colsum, _bf16(cst.view(-1))))
            self.__dict__["_ln_cache"] = hit
        return hit[1]

    def forward(self, x2d: torch.Tensor, residual: Optional[torch.Tensor] = None,
                ln_fold: Optional[LNInput] = None) -> torch.Tensor:
        """ln_fold: x2d is the raw residual stream, norm3 is folded into the GEGLU projection."""
        p = self.pack()
        if ln_fold is not None:
            wg, colsum, b1 = self.ln_pack(ln_fold.norm)
            h = ops.linear(x2d, wg, bias=b1, geglu=True, tile_n=GEGLU_TILE,
                           ln=ops.LNFold(ln_fold.stats, colsum, ln_fold.norm.eps))
        else:
            h = ops.linear(x2d, p["w1"], bias=p["b1"], geglu=True, tile_n=GEGLU_TILE)
        return ops.linear(h, p["w2"], bias=p["b2"], residual=residual)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, cross_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, cross_attention_dim=cross_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    @staticmethod
    def _ln(norm: nn.LayerNorm, x):
        return ops.layernorm(x, _bf16(norm.weight), _bf16(norm.bias), norm.eps)

    def _attend(self, attn: Attention, normed, h, ctx, kw):
        """h + attn(normed): the residual add is fused into the output projection when the processor takes a
        `residual` argument (ours do); a foreign processor gets the plain protocol + an add kernel."""
        if "residual" in accepted_params(attn.processor):
            return attn(normed, encoder_hidden_states=ctx, residual=h, **kw)
        return ops.add(h, attn(normed, encoder_hidden_states=ctx, **kw).contiguous())

    def can_fold_layernorm(self) -> bool:
        """Both processors take the raw stream + an LNInput (ours do; a foreign processor gets explicit LayerNorms)."""
        return FOLD_LAYERNORM and all({"ln_fold", "want_stats", "residual"} <= accepted_params(a.processor)
                                      for a in (self.attn1, self.attn2))

    def forward(self, h: torch.Tensor, encoder_hidden_states=None, cross_attention_kwargs=None,
                h_stats=None) -> torch.Tensor:
        """h_stats: ops.RowStats of h from the launch that produced it. With it the three LayerNorms never run as
        kernels: each is folded into the projection that consumes it, and every producing GEMM hands the row
        statistics of its output to the next one (diffusers BasicTransformerBlock)."""
        kw = dict(cross_attention_kwargs or {})
        b, s, c = h.shape
        if h_stats is not None and self.can_fold_layernorm():
            st = h_stats
            fold_ff = b * s <= FOLD_GEGLU_MAX_ROWS
            for attn, norm, ctx, want in ((self.attn1, self.norm1, None, True),
                                          (self.attn2, self.norm2, encoder_hidden_states, fold_ff)):
                if st is not None:
                    r = attn(h, encoder_hidden_states=ctx, residual=h, ln_fold=LNInput(norm, st), want_stats=want, **kw)
                else:  # the processor could not hand statistics on (foreign original processor): explicit LayerNorm
                    r = attn(self._ln(norm, h), encoder_hidden_states=ctx, residual=h, want_stats=want, **kw)
                h, st = r if want else (r, None)
            h2d = h.view(b * s, c)
            # The GEGLU projection is bound by its epilogue (erf GELU on 8C columns); measured in round 2
            # (profiles/r2_ln_fold.txt): with many rows the fold costs it more (+29 us at 32768 x 320) than the
            # LayerNorm kernel it removes (14 us), with few rows the launch is what counts.
            if st is not None and fold_ff:
                return self.ff(h2d, residual=h2d, ln_fold=LNInput(self.norm3, st)).view(b, s, c)
            return self.ff(self._ln(self.norm3, h).view(b * s, c), residual=h2d).view(b, s, c)
        h = self._attend(self.attn1, self._ln(self.norm1, h), h, None, kw)
        h = self._attend(self.attn2, self._ln(self.norm2, h), h, encoder_hidden_states, kw)
        n3 = self._ln(self.norm3, h)
        return self.ff(n3.view(b * s, c), residual=h.view(b * s, c)).view(b, s, c)


class Transformer2DModel(nn.Module):
    def __init__(self, dim: int, heads: int, cross_dim: int, groups: int = 32):
        super().__init__()
        self.groups = groups
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, encoder_hidden_states=None, cross_attention_kwargs=None, return_dict=False, out_film=None):
        """out_film = (scale, shift) fp32 [B, C]: camera FiLM of the enclosing block's output, applied in proj_out's
        epilogue (after the residual)."""
        xa = nhwc_view(x)
        n, hh, ww, c = xa.shape
        res2d = xa.reshape(n * hh * ww, c)
        h = ops.groupnorm(xa, _bf16(self.norm.weight), _bf16(self.norm.bias), self.groups, self.norm.eps, silu=False)
        fold = len(self.transformer_blocks) == 1 and self.transformer_blocks[0].can_fold_layernorm()
        h = ops.linear(h.view(n * hh * ww, c), _bf16(self.proj_in.weight), bias=_bf16(self.proj_in.bias),
                       want_stats=fold)
        h, h_stats = h if fold else (h, None)
        h = h.view(n, hh * ww, c)
        for blk in self.transformer_blocks:
            h = blk(h, encoder_hidden_states, cross_attention_kwargs, h_stats=h_stats)
        out = ops.linear(h.view(n * hh * ww, c), _bf16(self.proj_out.weight), bias=_bf16(self.proj_out.bias),
                         residual=res2d, rows_per_group=hh * ww if out_film is not None else 0, film=out_film)
        return (nchw_shape(out.view(n, hh, ww, c)),)


class _PackedConv(PackedModule):
    def _pack_params(self):
        return [self.conv.weight, self.conv.bias]

    def _build_pack(self):
        return dict(w=conv3x3_weight(self.conv.weight), b=_bf16(self.conv.bias))


class Downsample2D(_PackedConv):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        p = self.pack()
        return nchw_shape(ops.conv3x3(nhwc_view(x), p["w"], bias=p["b"], stride=2))


class Upsample2D(_PackedConv):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x, film=None):
        """film = (scale, shift) fp32 [B, C]: camera FiLM of the block output, applied in the conv epilogue."""
        p = self.pack()
        return nchw_shape(ops.conv3x3(ops.upsample2x(nhwc_view(x)), p["w"], bias=p["b"], film=film))


class DownBlock(nn.Module):
    """CrossAttnDownBlock2D / DownBlock2D."""

    def __init__(self, cin, cout, heads, cross_dim, has_attn, add_down, temb_dim=1280):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim) for i in range(2)])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim) for _ in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None
        self.has_attn = has_attn

    def forward(self, h, temb, encoder_hidden_states=None, cross_attention_kwargs=None, temb_projs=None):
        states = ()
        for i, res in enumerate(self.resnets):
            h = res(h, temb, temb_proj=None if temb_projs is None else temb_projs[id(res)])
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

    def forward(self, h, temb, encoder_hidden_states=None, cross_attention_kwargs=None, temb_projs=None):
        tp = (lambda r: None) if temb_projs is None else (lambda r: temb_projs[id(r)])
        h = self.resnets[0](h, temb, temb_proj=tp(self.resnets[0]))
        h = self.attentions[0](h, encoder_hidden_states, cross_attention_kwargs)[0]
        return self.resnets[1](h, temb, temb_proj=tp(self.resnets[1]))


class UpBlock(nn.Module):
    """CrossAttnUpBlock2D / UpBlock2D. The torch.cat([h, skip]) of diffusers is fused into the resnet."""

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
        # camera FiLM of this block's output (MultiViewUNet sets it per forward): (scale, shift) fp32 [B, C]; when the
        # block's last op can apply it in its epilogue, `film_applied` tells the forward hook not to do it again
        self.out_film = None
        self.film_applied = False

    def forward(self, h, skips, temb, encoder_hidden_states=None, cross_attention_kwargs=None, temb_projs=None):
        film, self.film_applied = self.out_film, False
        n_res = len(self.resnets)
        for i, res in enumerate(self.resnets):
            skip, skips = skips[-1], skips[:-1]
            h = res(h, temb, skip=skip, temb_proj=None if temb_projs is None else temb_projs[id(res)])
            if self.has_attn:
                last = film is not None and i == n_res - 1 and self.upsamplers is None
                h = self.attentions[i](h, encoder_hidden_states, cross_attention_kwargs,
                                       out_film=film if last else None)[0]
                self.film_applied = self.film_applied or last
        if self.upsamplers is not None:
            h = self.upsamplers[0](h, film=film)
            self.film_applied = film is not None
        return h


class UNetOut(tuple):
    @property
    def sample(self):
        return self[0]


class UNet2DConditionModel(PackedModule):
    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(SD21_CONFIG)
        cfg.update(overrides)
        self.config = _Config(cfg)
        ch, heads, cross = cfg["block_out_channels"], cfg["attention_head_dim"], cfg["cross_attention_dim"]
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
        self.input_film = None  # (mod fp32 [V,8], strength): camera FiLM on the input latents, fused into conv_in
        self.temb_rows = None   # fp32 [sum of resnet out_channels]: the current row of a DenoiseSession's per-schedule table

    @property
    def device(self):
        return self.conv_in.weight.device

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def enable_gradient_checkpointing(self):  # accepted for API compatibility (inference-only kernels)
        pass

    def _resnets(self):
        return [m for m in self.modules() if isinstance(m, ResnetBlock2D)]

    def _pack_params(self):
        ps = [self.conv_in.weight, self.conv_in.bias, self.conv_out.weight, self.conv_out.bias,
              self.conv_norm_out.weight, self.conv_norm_out.bias]
        for r in self._resnets():
            ps += [r.time_emb_proj.weight, r.time_emb_proj.bias]
        return ps

    def _build_pack(self):
        res = self._resnets()
        offs, o = {}, 0
        for r in res:
            offs[id(r)] = (o, o + r.out_channels)
            o += r.out_channels
        return dict(
            w_in=_bf16(self.conv_in.weight.permute(0, 2, 3, 1)), b_in=_bf16(self.conv_in.bias),
            w_out=_bf16(self.conv_out.weight.permute(0, 2, 3, 1)), b_out=_bf16(self.conv_out.bias),
            # conv_out as a 32-column tcgen05 conv: the 4 weight rows [(ky,kx,c) order] zero-padded to the narrowest tile
            w_out32=self._pad_rows(_bf16(self.conv_out.weight.permute(0, 2, 3, 1)).reshape(self.conv_out.out_channels, -1), 32),
            b_out32=self._pad_rows(_bf16(self.conv_out.bias), 32),
            g_out=_bf16(self.conv_norm_out.weight), bn_out=_bf16(self.conv_norm_out.bias),
            # all 22 time_emb_proj layers as one skinny GEMM
            tw=_bf16(torch.cat([r.time_emb_proj.weight for r in res], 0)),
            tb=_bf16(torch.cat([r.time_emb_proj.bias for r in res], 0)), toffs=offs,
        )

    @staticmethod
    def _pad_rows(t: torch.Tensor, rows: int) -> torch.Tensor:
        out = torch.zeros((rows,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
        out[: t.shape[0]].copy_(t)
        return out

    def timestep_rows(self, t: torch.Tensor, bsz: int):
        """diffusers Timesteps + TimestepEmbedding, then all 22 time_emb_proj(SiLU(temb)) as one skinny GEMM:
        (temb fp32 [bsz, 1280], rows fp32 [bsz, sum of resnet out_channels])."""
        p = self.pack()
        temb = self.time_embedding(ops.timestep_embedding(t, bsz, self.config.block_out_channels[0]))
        return temb, _small_linear_any_m(temb, p["tw"], p["tb"], silu_in=True)

    def forward(self, sample, timestep, encoder_hidden_states, return_dict: bool = True, timestep_cond=None,
                cross_attention_kwargs: Optional[Dict[str, Any]] = None, added_cond_kwargs=None):
        """sample: fp32 (or bf16) NCHW latents [B,4,H,W]; timestep: python number or tensor (scalar or [B]);
        encoder_hidden_states: [B,77,1024]. Returns UNetOut((fp32 NCHW [B,4,H,W],))."""
        if not sample.is_cuda:
            raise ValueError("mvd_b200 runs on CUDA tensors only")
        p = self.pack()
        dev = sample.device
        bsz = sample.shape[0]
        if torch.is_tensor(timestep):
            t = timestep.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        else:
            t = torch.full((1,), float(timestep), device=dev, dtype=torch.float32)
        rows = getattr(self, "temb_rows", None)
        if rows is not None:
            # per-schedule table (DenoiseSession): the current step's time_emb_proj outputs, identical for every
            # sample of the batch (stride-0 rows), no timestep arithmetic in the step
            temb, tp_all = None, rows.view(1, -1).expand(bsz, -1)
        else:
            temb, tp_all = self.timestep_rows(t, bsz)
        temb_projs = {k: tp_all[:, a:b] for k, (a, b) in p["toffs"].items()}

        lat = sample.float().contiguous() if sample.dtype != torch.float32 or not sample.is_contiguous() else sample
        mod, strength = self.input_film if self.input_film is not None else (None, 1.0)
        h = nchw_shape(ops.conv_in(lat, p["w_in"], p["b_in"], n_img=bsz, mod=mod, strength=strength))

        ctx = encoder_hidden_states
        skips: Tuple[torch.Tensor, ...] = (h,)
        for blk in self.down_blocks:
            h, st = blk(h, temb, ctx, cross_attention_kwargs, temb_projs=temb_projs)
            skips += st
        h = self.mid_block(h, temb, ctx, cross_attention_kwargs, temb_projs=temb_projs)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            h = blk(h, skips[-n:], temb, ctx, cross_attention_kwargs, temb_projs=temb_projs)
            skips = skips[:-n]
        hn = ops.groupnorm(nhwc_view(h), p["g_out"], p["bn_out"], self.config.norm_num_groups, self.config.norm_eps,
                           silu=True)
        if CONV_OUT_TENSOR_CORES and self.conv_out.out_channels == 4 and hn.shape[-1] % 64 == 0:
            # 8 x 64 x 64 x 320 -> 4: 120 us as a CUDA-core kernel, 16 us as a 32-column implicit GEMM + the 4-channel tail
            out = ops.head4_to_nchw(ops.conv3x3(hn, p["w_out32"], bias=p["b_out32"]))
        else:
            out = ops.conv_out(hn, p["w_out"], p["b_out"])
        return UNetOut((out,))


def tiny_config() -> dict:
    """Structurally identical small UNet (same block types, head_dim 64, channels multiple of 64) for tests."""
    return dict(block_out_channels=(64, 128, 128, 128), attention_head_dim=(1, 2, 2, 2), cross_attention_dim=64)
