"""Drop-in for the reference's `src/models/attention.py`: same class / factory names, same parameters
(state-dict keys), same attention-processor protocol — the arithmetic runs in libmvd_b200.so.

    out = original_processor(h) + ref_scale * to_out_ref(SDPA(to_q_ref(h), to_k_ref(r), to_v_ref(r)))
    r   = normalise(ref_hidden_states[name])            (reference attention.py:95-103)

B200 mapping of one processor call (reference attention.py:48-188):
  1. ONE tcgen05 GEMM projects h to [q | k | v | q_ref] (self) or [q | q_ref] (text cross-attention).
  2. K/V of the normalised reference features are STEP-INVARIANT (the reference features come from a frozen
     UNet at t = 0): refnorm + one GEMM, cached per reference tensor. K/V of the text likewise.
  3. Two flash-attention launches read their heads in place and write O_orig | O_ref side by side.
  4. ONE GEMM with K = 2C applies [W_out | ref_scale * W_out_ref], adds both biases and the transformer
     block's residual in its epilogue.
`S_kv` of the reference branch is arbitrary: a 3-D `[B, S_kv, C]` reference (e.g. all N views' tokens
concatenated, plus any injected rows) goes through the same kernels (north-star "cross-view" mode).
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .unet import BF16, AttnProcessor2_0, LNInput, _bf16, _versions, fold_layernorm, nhwc_view


def log_debug(file_path, message):  # reference src/utils.py:25-34; callers here never build tensor f-strings
    return None


# Run the reference-attention branch concurrently with the original branch (second CUDA stream, fork/join by events).
# MVD_OVERLAP_BRANCHES=0 launches them back to back on the current stream instead.
OVERLAP_BRANCHES = os.environ.get("MVD_OVERLAP_BRANCHES", "1") != "0"
_branch_streams: Dict[int, "torch.cuda.Stream"] = {}


_sm_counts: Dict[int, int] = {}


def _sm_count(device: torch.device) -> int:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    n = _sm_counts.get(idx)
    if n is None:
        n = _sm_counts[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return n


def _branch_stream(device: torch.device) -> "torch.cuda.Stream":
    idx = device.index if device.index is not None else torch.cuda.current_device()
    s = _branch_streams.get(idx)
    if s is None:
        s = _branch_streams[idx] = torch.cuda.Stream(device=idx)
    return s


class SharedReference:
    """A 3-D reference [B, S_kv, C] whose B batch entries are IDENTICAL (cross-view mode, BASELINE configs[3]: every
    sample attends over the tokens of all views), held as ONE copy `tokens` [1, S_kv, C] + the batch size it stands
    for. The reference normalisation (attention.py:95-103, statistics over (batch, sequence) of the raw tensor) is
    computed for the replicated tensor without materialising it, K/V are projected once and every sample reads them
    through a zero batch stride."""

    __slots__ = ("tokens", "replication")

    def __init__(self, tokens: torch.Tensor, replication: int):
        if tokens.dim() != 3 or tokens.shape[0] != 1:
            raise ValueError("SharedReference holds one copy: tokens must be [1, S_kv, C]")
        self.tokens, self.replication = tokens, int(replication)


class ImageCrossAttentionProcessor(nn.Module):
    def __init__(self, name: str, query_dim: int, heads: int, dim_head: int = 64, dropout: float = 0.0,
                 img_ref_scale: float = 0.3):
        super().__init__()
        self.name = name
        self.heads = heads
        self.dim_head = dim_head
        self.inner_dim = heads * dim_head
        self.query_dim = query_dim
        self.original_processor = None
        self.to_q_ref = nn.Linear(query_dim, self.inner_dim, bias=False)
        self.to_k_ref = nn.Linear(query_dim, self.inner_dim, bias=False)
        self.to_v_ref = nn.Linear(query_dim, self.inner_dim, bias=False)
        self.ref_ln = nn.LayerNorm(self.inner_dim)  # registered, trainable, unused — as in the reference (:37,160)
        self.feature_adapter = None
        self.to_out_ref = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=True), nn.Dropout(dropout)])
        self.ref_scale_val = img_ref_scale

    # ---- packed weights -----------------------------------------------------------------------------------
    def _pack(self, attn):
        own = [self.to_q_ref.weight, self.to_k_ref.weight, self.to_v_ref.weight, self.to_out_ref[0].weight,
               self.to_out_ref[0].bias]
        theirs = [attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight, attn.to_out[0].bias]
        key = (_versions(*own, *theirs), float(self.ref_scale_val))
        cached = self.__dict__.get("_pack_cache")
        if cached is not None and cached[0] == key:
            return cached[1]
        s = float(self.ref_scale_val)
        with torch.no_grad():
            fused_native = isinstance(self.original_processor, AttnProcessor2_0)
            p: Dict[str, Any] = dict(fused=fused_native)
            p["wkv_ref"] = _bf16(torch.cat([self.to_k_ref.weight, self.to_v_ref.weight], 0))
            if fused_native:
                is_cross = attn.to_k.weight.shape[1] != attn.to_q.weight.shape[1] or getattr(attn, "is_cross", False)
                p["is_cross"] = is_cross
                if is_cross:
                    p["w_in"] = _bf16(torch.cat([attn.to_q.weight, self.to_q_ref.weight], 0))  # [q | q_ref]
                else:
                    p["w_in"] = _bf16(torch.cat([attn.to_q.weight, attn.to_k.weight, attn.to_v.weight,
                                                 self.to_q_ref.weight], 0))  # [q | k | v | q_ref]
                p["w_out"] = _bf16(torch.cat([attn.to_out[0].weight.float(), s * self.to_out_ref[0].weight.float()], 1))
                p["b_out"] = _bf16(attn.to_out[0].bias.float() + s * self.to_out_ref[0].bias.float())
            else:
                p["wq_ref"] = _bf16(self.to_q_ref.weight)
                p["w_out"] = _bf16(s * self.to_out_ref[0].weight.float())
                p["b_out"] = _bf16(s * self.to_out_ref[0].bias.float())
        self.__dict__["_pack_cache"] = (key, p)
        self.__dict__["_ref_cache"] = None
        return p

    def _ln_pack(self, attn, norm: nn.LayerNorm, pk):
        """[q | k | v | q_ref] (or [q | q_ref]) with the block's LayerNorm folded in (unet.fold_layernorm)."""
        key = (id(pk), _versions(norm.weight, norm.bias))
        hit = self.__dict__.get("_ln_cache")
        if hit is None or hit[0] != key or hit[2] is not pk:
            with torch.no_grad():
                hit = (key, fold_layernorm(pk["w_in"], norm), pk)
            self.__dict__["_ln_cache"] = hit
        return hit[1]

    # ---- step-invariant reference K/V ---------------------------------------------------------------------
    def _reference_kv(self, ref: torch.Tensor, pk, replication: int = 1) -> torch.Tensor:
        """[B_ref * S_kv, 2C] = [to_k_ref(r) | to_v_ref(r)], r = normalised reference (attention.py:95-132)."""
        key = (ref.data_ptr(), ref._version, tuple(ref.shape), tuple(ref.stride()), ref.dtype, replication)
        cached = self.__dict__.get("_ref_cache")
        if cached is not None and cached[0] == key:
            return cached[1]
        if not ref.is_cuda:
            raise ValueError("reference features must be CUDA tensors")
        if ref.dim() == 4:  # [B,C,H,W]: statistics per pixel over (batch, channel); tokens = NHWC rows
            tok = nhwc_view(ref if ref.dtype in (BF16, torch.float32) else ref.float())
            b, hh, ww, c = tok.shape
            tok = tok.reshape(b, hh * ww, c)
            per_pixel = True
        elif ref.dim() == 3:  # [B,S,C]: statistics per channel over (batch, sequence)
            tok = ref
            if tok.dtype != BF16:
                tok = ops.cast_bf16(tok.float().contiguous())
            tok = tok.contiguous()
            per_pixel = False
        else:
            raise ValueError(f"reference features must be 3-D or 4-D, got {ref.dim()}-D")
        if tok.shape[-1] != self.query_dim:
            raise ValueError(f"{self.name}: reference has {tok.shape[-1]} channels, expected {self.query_dim}")
        if replication != 1 and per_pixel:
            raise ValueError("a shared (replicated) reference must be 3-D")
        normed = ops.refnorm(tok.contiguous(), per_pixel=per_pixel, replication=replication)
        b, s, c = normed.shape
        kv = ops.linear(normed.view(b * s, c), pk["wkv_ref"])
        self.__dict__["_ref_cache"] = (key, kv, ref)  # hold `ref` so its storage cannot be recycled under the key
        return kv

    def _gather_batch(self, kv_ref: torch.Tensor, b_ref: int, index: torch.Tensor) -> torch.Tensor:
        key = (kv_ref.data_ptr(), index.data_ptr(), index._version)
        cached = self.__dict__.get("_gather_cache")
        if cached is not None and cached[0] == key:
            return cached[1]
        picked = kv_ref.view(b_ref, -1, kv_ref.shape[1]).index_select(0, index)
        picked = picked.reshape(-1, kv_ref.shape[1]).contiguous()
        self.__dict__["_gather_cache"] = (key, picked, kv_ref, index)
        return picked

    # ---- the processor protocol ---------------------------------------------------------------------------
    def __call__(self, attn: Any, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, temb: Optional[torch.Tensor] = None,
                 ref_hidden_states: Optional[Dict[str, torch.Tensor]] = None,
                 residual: Optional[torch.Tensor] = None, ref_batch_index: Optional[torch.Tensor] = None,
                 ln_fold: Optional[LNInput] = None, want_stats: bool = False, *args, **kwargs):
        """Extensions of the diffusers protocol used by mvd_b200's own transformer block (a stock caller never
        passes them): `residual` (added in the output projection's epilogue), `ln_fold` (hidden_states is the RAW
        residual stream and the block's LayerNorm is folded into the input projection), `want_stats` (return
        (out, ops.RowStats) so the next LayerNorm can be folded too)."""
        kwargs.pop("debug_log_file_path", None)
        native = isinstance(self.original_processor, AttnProcessor2_0)
        if ref_hidden_states is None or self.name not in ref_hidden_states:
            # reference attention.py:72-81: silently fall through to the original processor
            if native:
                return self.original_processor(attn, hidden_states, encoder_hidden_states, attention_mask, temb=temb,
                                               residual=residual, ln_fold=ln_fold, want_stats=want_stats)
            if ln_fold is not None:
                hidden_states = ops.layernorm(hidden_states, _bf16(ln_fold.norm.weight), _bf16(ln_fold.norm.bias),
                                              ln_fold.norm.eps)
            out = self.original_processor(attn, hidden_states, encoder_hidden_states, attention_mask, temb=temb,
                                          *args, **kwargs)
            out = out if residual is None else ops.add(out.contiguous(), residual.contiguous())
            return (out, None) if want_stats else out
        if attention_mask is not None:
            raise NotImplementedError("attention masks are not on MVD's hot path")
        if hidden_states.dim() != 3:
            raise ValueError("hidden_states must be [B, HW, C] (the 4-D branch of the reference is dead code)")
        if not hidden_states.is_cuda or hidden_states.dtype != BF16:
            raise ValueError("hidden_states must be a CUDA bf16 tensor (no CPU / fp32 fallback path)")
        pk = self._pack(attn)
        b, s, c = hidden_states.shape
        hs2d = hidden_states.reshape(b * s, c)
        ref_t = ref_hidden_states[self.name]
        shared = isinstance(ref_t, SharedReference)
        if shared:
            if ref_t.replication < b:
                raise ValueError(f"{self.name}: the shared reference stands for {ref_t.replication} samples, query "
                                 f"batch is {b}")
            kv_ref = self._reference_kv(ref_t.tokens, pk, replication=ref_t.replication)
        else:
            kv_ref = self._reference_kv(ref_t, pk)
        if ref_batch_index is not None and not shared:
            # view-sharded execution (mvd_b200/dist.py): K/V were normalised + projected over the FULL reference
            # batch; this rank attends with the rows of its own samples
            kv_ref = self._gather_batch(kv_ref, ref_t.shape[0], ref_batch_index)
        rows = kv_ref.shape[0]
        if shared:  # every sample reads the one K/V copy (zero batch stride in the attention kernel's tensor maps)
            k_ref = kv_ref[:, :c].unsqueeze(0).expand(b, rows, c)
            v_ref = kv_ref[:, c:].unsqueeze(0).expand(b, rows, c)
        else:
            if rows % b:
                raise ValueError(f"{self.name}: {rows} reference tokens do not split over query batch {b}")
            # reference attention.py:130,132: key.view(batch_size, -1, heads, dim_head) — a flat re-view by the QUERY
            # batch (with CFG and an un-repeated reference each sample sees a different half of the tokens)
            k_ref = kv_ref[:, :c].view(b, rows // b, c)
            v_ref = kv_ref[:, c:].view(b, rows // b, c)
        res2d = residual.reshape(b * s, c) if residual is not None else None
        scale = self.dim_head ** -0.5

        if pk["fused"]:
            cat = torch.empty((b, s, 2 * c), device=hidden_states.device, dtype=BF16)
            if ln_fold is not None:
                wg, colsum, cst = self._ln_pack(attn, ln_fold.norm, pk)
                proj = ops.linear(hs2d, wg, row_group_bias=cst, rows_per_group=b * s,
                                  ln=ops.LNFold(ln_fold.stats, colsum, ln_fold.norm.eps))
            else:
                proj = ops.linear(hs2d, pk["w_in"])
            if pk["is_cross"]:
                proj = proj.view(b, s, 2 * c)
                q, q_ref = proj[:, :, :c], proj[:, :, c:]
                kv = attn.context_kv(encoder_hidden_states)
                k, v = kv[:, :, :c], kv[:, :, c:]
            else:
                proj = proj.view(b, s, 4 * c)
                q, k, v, q_ref = (proj[:, :, i * c:(i + 1) * c] for i in range(4))
            # The two branches are independent until the fused out-projection. A launch that owns the machine
            # (persistent CTAs) is simply followed by its sibling; otherwise each launch is a non-integral number of
            # one-CTA-per-SM waves and the reference branch is forked onto a second stream to fill the tail of the
            # first (inside a captured step this becomes two parallel graph branches).
            sms = _sm_count(hidden_states.device)
            units = ops.attention_units(b, self.heads, s)
            big = units >= sms
            if not pk["is_cross"] and (ops.attention_is_persistent(b, self.heads, s, k.shape[1], sms) or
                                       ops.attention_is_persistent(b, self.heads, s, k_ref.shape[1], sms)):
                # each launch is a persistent grid of one CTA per SM that splits its own last wave along S_kv: nothing
                # is left for a sibling to fill, so the two branches simply follow each other on this stream
                ops.attention(q, k, v, self.heads, scale, out=cat[:, :, :c])
                ops.attention(q_ref, k_ref, v_ref, self.heads, scale, out=cat[:, :, c:])
            elif OVERLAP_BRANCHES:
                main = torch.cuda.current_stream()
                side = _branch_stream(hidden_states.device)
                fork, join = torch.cuda.Event(), torch.cuda.Event()
                fork.record(main)
                # Who splits its last partial wave along S_kv (ops.attention): a launch of >= 1 wave of units shares
                # the machine's last wave with its sibling — the side launch goes first and finishes first, so only
                # the main one splits, and only its share of the COMBINED last wave (co_units). Small launches (a
                # view-sharded rank: fewer units than SMs) both split. Next to the tiny text attention the reference
                # branch simply owns the machine.
                if pk["is_cross"]:
                    side_kw, main_kw = dict(split_tail=True), dict(split_tail=False)
                elif big:
                    side_kw, main_kw = dict(split_tail=False), dict(co_units=units)
                else:
                    side_kw, main_kw = dict(co_units=units), dict(co_units=units)
                with torch.cuda.stream(side):
                    side.wait_event(fork)
                    ops.attention(q_ref, k_ref, v_ref, self.heads, scale, out=cat[:, :, c:], **side_kw)
                    join.record(side)
                ops.attention(q, k, v, self.heads, scale, out=cat[:, :, :c], **main_kw)
                main.wait_event(join)
            else:
                ops.attention(q, k, v, self.heads, scale, out=cat[:, :, :c])
                ops.attention(q_ref, k_ref, v_ref, self.heads, scale, out=cat[:, :, c:])
            out = ops.linear(cat.view(b * s, 2 * c), pk["w_out"], bias=pk["b_out"], residual=res2d,
                             want_stats=want_stats)
            if want_stats:
                return out[0].view(b, s, c), out[1]
            return out.view(b, s, c)

        # foreign original processor (e.g. a stock diffusers processor): run it as is, add our branch on top
        if ln_fold is not None:
            hidden_states = ops.layernorm(hidden_states, _bf16(ln_fold.norm.weight), _bf16(ln_fold.norm.bias),
                                          ln_fold.norm.eps)
            hs2d = hidden_states.reshape(b * s, c)
        original = self.original_processor(attn, hidden_states, encoder_hidden_states, attention_mask, temb=temb,
                                           *args, **kwargs)
        base = original.reshape(b * s, c).contiguous()
        if res2d is not None:
            base = ops.add(base, res2d.contiguous())
        q_ref = ops.linear(hs2d, pk["wq_ref"]).view(b, s, c)
        o = ops.attention(q_ref, k_ref, v_ref, self.heads, scale)
        out = ops.linear(o.view(b * s, c), pk["w_out"], bias=pk["b_out"], residual=base, want_stats=want_stats)
        if want_stats:
            return out[0].view(b, s, c), out[1]
        return out.view(b, s, c)

    def _adapt_reference_features(self, reference_states, target_dim):
        """reference attention.py:190-197 (kept for API parity): NCHW -> [B, HW, C] view."""
        if reference_states.ndim == 4:
            t = nhwc_view(reference_states)
            return t.reshape(t.shape[0], -1, t.shape[-1])
        return reference_states

    def load_original_weights(self, attn_module):
        """reference attention.py:199-246: q/out copied; k/v copied when shapes agree (self-attention), else a
        transposed leading slice of the text-attention weights, or zero-padded columns."""
        with torch.no_grad():
            self.to_q_ref.weight.copy_(attn_module.to_q.weight)
            self.to_out_ref[0].weight.copy_(attn_module.to_out[0].weight)
            self.to_out_ref[0].bias.copy_(attn_module.to_out[0].bias)
            for dst, src in ((self.to_k_ref.weight, attn_module.to_k.weight),
                             (self.to_v_ref.weight, attn_module.to_v.weight)):
                d_out, d_in = dst.shape
                s_out, s_in = src.shape
                if (d_out, d_in) == (s_out, s_in):
                    dst.copy_(src)
                elif d_in >= s_in:
                    dst[:, :s_in].copy_(src[: min(d_out, s_out), :])
                    if d_in > s_in:
                        dst[:, s_in:].zero_()
                else:
                    dst.copy_(src[: min(d_out, s_out), :d_in].t())


def get_attention_processor_for_module(name, attn_module, img_ref_scale=0.3):
    """reference attention.py:248-265."""
    query_dim = attn_module.to_q.in_features
    heads = attn_module.heads
    processor = ImageCrossAttentionProcessor(name=name, query_dim=query_dim, heads=heads,
                                             dim_head=attn_module.to_q.out_features // heads,
                                             img_ref_scale=img_ref_scale)
    processor.original_processor = attn_module.processor
    w = attn_module.to_q.weight
    processor.to(device=w.device, dtype=w.dtype)
    processor.load_original_weights(attn_module)
    return processor
