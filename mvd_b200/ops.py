"""Thin torch-tensor wrappers over the C ABI (include/mvd_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream. Every function validates
device/dtype/layout, passes raw pointers to libmvd_b200.so and raises on any error. No op has a PyTorch
or CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from ._lib import check, lib

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=BF16):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (mvd_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")


def _rows2d(t: torch.Tensor, name: str):
    """Return (ptr, ld, rows, cols) of a 2-D row-strided view."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} stride {t.stride()}")
    return t.data_ptr(), t.stride(0), t.shape[0], t.shape[1]


# split-K of under-filled GEMM / conv launches (csrc/gemm.cu); MVD_SPLIT_K=0 switches it off for A/B runs
SPLIT_K = os.environ.get("MVD_SPLIT_K", "1") != "0"
# launches of at least one wave of attention units run as persistent CTAs (csrc/attn.cu); they own the machine, so the
# two attention branches of an adapter block are then launched back to back instead of on two streams
ATTN_PERSIST = os.environ.get("MVD_ATTN_PERSIST", "1") != "0"


class _GemmExtras(ctypes.Structure):  # mirrors mvd_gemm_extras (include/mvd_b200.h)
    _fields_ = [("ln_stats", ctypes.c_void_p), ("ln_colsum", ctypes.c_void_p), ("ln_parts", ctypes.c_int),
                ("ln_eps", ctypes.c_float), ("stats_out", ctypes.c_void_p), ("stats_parts", ctypes.c_int),
                ("film_scale", ctypes.c_void_p), ("film_shift", ctypes.c_void_p), ("film_ld", ctypes.c_int),
                ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64)]


class RowStats:
    """Partial (sum, sum of squares) of every row of a [M, C] bf16 activation, one pair per column tile of the launch
    that produced it: what the LayerNorm folded into the NEXT GEMM needs (buf: fp32 [M, parts, 2])."""

    __slots__ = ("buf", "parts", "channels")

    def __init__(self, buf: torch.Tensor, parts: int, channels: int):
        self.buf, self.parts, self.channels = buf, parts, channels


class LNFold:
    """LayerNorm(x) @ W^T folded into one GEMM on the raw x: `stats` of x, `colsum[n]` = sum_k of the gamma-scaled bf16
    weight row n (fp32); the caller's weight carries gamma and its bias carries W.beta."""

    __slots__ = ("stats", "colsum", "eps")

    def __init__(self, stats: RowStats, colsum: torch.Tensor, eps: float):
        self.stats, self.colsum, self.eps = stats, colsum, float(eps)


_PLAN_CACHE: dict = {}


def linear_column_tiles(M: int, N: int, K: int, geglu: bool = False, tile_n: int = 0) -> int:
    """Column-tile count of the mvd_linear launch for this problem (mvd_gemm_plan): the `parts` of its RowStats."""
    key = (M, N, K, geglu, tile_n)
    got = _PLAN_CACHE.get(key)
    if got is None:
        bn = ctypes.c_int()
        check(lib().mvd_gemm_plan(1, 1, M, K, N, 1, 1, int(geglu), tile_n, ctypes.byref(bn), None, None),
              "mvd_gemm_plan")
        got = (N + bn.value - 1) // bn.value
        _PLAN_CACHE[key] = got
    return got


_GEMM_WS_FLOATS = 16 << 20  # 64 MiB of stream-K scratch (fp32 partial tiles) per (device, stream)


def _extras(ln: Optional["LNFold"], stats: Optional["RowStats"], film, N: int, M: int, device=None):
    ex = _GemmExtras()
    keep = []
    if device is not None and SPLIT_K:
        ws = _workspace(device, _GEMM_WS_FLOATS, "gemm")
        ex.workspace, ex.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    if ln is not None:
        _contig(ln.stats.buf, "ln.stats", F32)
        _contig(ln.colsum, "ln.colsum", F32)
        if ln.colsum.numel() != N or ln.stats.buf.shape[0] != M:
            raise ValueError("LayerNorm fold: colsum must be [N] and stats cover the M rows of a")
        ex.ln_stats, ex.ln_colsum = ln.stats.buf.data_ptr(), ln.colsum.data_ptr()
        ex.ln_parts, ex.ln_eps = ln.stats.parts, ln.eps
    if stats is not None:
        ex.stats_out, ex.stats_parts = stats.buf.data_ptr(), stats.parts
    if film is not None:
        scale, shift = film
        _req(scale, "film scale", F32)
        _req(shift, "film shift", F32)
        if scale.shape != shift.shape or scale.dim() != 2 or scale.shape[1] != N or scale.stride(1) != 1 or \
                shift.stride() != scale.stride():
            raise ValueError("film = (scale, shift): two fp32 [groups, N] tensors of identical layout")
        ex.film_scale, ex.film_shift, ex.film_ld = scale.data_ptr(), shift.data_ptr(), scale.stride(0)
    return ex, keep


def linear(
    a: torch.Tensor,
    w: torch.Tensor,
    bias: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    a2: Optional[torch.Tensor] = None,
    geglu: bool = False,
    row_group_bias: Optional[torch.Tensor] = None,
    rows_per_group: int = 0,
    out: Optional[torch.Tensor] = None,
    tile_n: int = 0,
    ln: Optional[LNFold] = None,
    want_stats: bool = False,
    film=None,
):
    """out[M,N] = [a|a2] @ w^T (+bias) (+row_group_bias[row // rows_per_group]) (+residual); optional GEGLU.
    ln: LayerNorm of a's rows folded in (see LNFold); want_stats: also return the RowStats of `out` for the LayerNorm
    that follows -> (out, RowStats); film = (scale, shift) fp32 [groups, N]: out = out * scale[g] + shift[g] with
    g = row // rows_per_group (camera FiLM on a block output)."""
    _req(a, "a")
    _req(w, "w")
    pa, lda, M, k1 = _rows2d(a, "a")
    pw, ldw, N, K = _rows2d(w, "w")
    pa2, lda2, k2 = None, 0, 0
    if a2 is not None:
        _req(a2, "a2")
        pa2, lda2, M2, k2 = _rows2d(a2, "a2")
        if M2 != M:
            raise ValueError("a and a2 must have the same number of rows")
    if k1 + k2 != K:
        raise ValueError(f"inner dimensions differ: a has {k1}+{k2}, w has {K}")
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=BF16)
    _req(out, "out")
    po, ldo, Mo, No = _rows2d(out, "out")
    if (Mo, No) != (M, n_out):
        raise ValueError(f"out must be [{M},{n_out}], got {tuple(out.shape)}")
    pr, ldr = None, 0
    if residual is not None:
        _req(residual, "residual")
        pr, ldr, Mr, Nr = _rows2d(residual, "residual")
        if (Mr, Nr) != (M, n_out):
            raise ValueError("residual shape mismatch")
    if bias is not None:
        _req(bias, "bias")
        if bias.numel() != N or not bias.is_contiguous():
            raise ValueError("bias must be a contiguous [N] tensor")
    pg, ldg = None, 0
    if row_group_bias is not None:
        _req(row_group_bias, "row_group_bias", torch.float32)
        pg, ldg, _, Ng = _rows2d(row_group_bias, "row_group_bias")
        if Ng != N or rows_per_group <= 0:
            raise ValueError("row_group_bias must be [groups, N] with rows_per_group > 0")
    stats = None
    if want_stats:
        parts = linear_column_tiles(M, N, K, geglu, tile_n)
        stats = RowStats(torch.empty((M, parts, 2), device=a.device, dtype=F32), parts, n_out)
    if film is not None and rows_per_group <= 0:
        raise ValueError("film needs rows_per_group > 0")
    ex, _keep = _extras(ln, stats, film, N, M, a.device)
    check(
        lib().mvd_linear_ex_bf16(pa, lda, k1, pa2, lda2, k2, pw, ldw, _p(bias), pg, ldg, rows_per_group, pr, ldr, po,
                                 ldo, M, N, int(geglu), tile_n, ctypes.byref(ex), _stream()),
        "mvd_linear_ex_bf16",
    )
    return (out, stats) if want_stats else out


def conv3x3(
    x: torch.Tensor,
    w: torch.Tensor,
    bias: Optional[torch.Tensor] = None,
    img_bias: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    x2: Optional[torch.Tensor] = None,
    stride: int = 1,
    out: Optional[torch.Tensor] = None,
    tile_n: int = 0,
    film=None,
) -> torch.Tensor:
    """NHWC 3x3 conv, padding 1. x: [N,H,W,C1] (x2: [N,H,W,C2] concatenated after x); w: [Cout, 9*(C1+C2)]
    in (ky, kx, c) order; img_bias: fp32 [N, Cout]; residual/out: [N,Ho,Wo,Cout]."""
    _req(x, "x")
    _req(w, "w")
    if x.dim() != 4 or not x.is_contiguous():
        raise ValueError("x must be a contiguous [N,H,W,C] tensor")
    n, h, wd, c1 = x.shape
    c2 = 0
    if x2 is not None:
        _req(x2, "x2")
        if x2.dim() != 4 or not x2.is_contiguous() or x2.shape[:3] != x.shape[:3]:
            raise ValueError("x2 must be contiguous [N,H,W,C2] matching x")
        c2 = x2.shape[3]
    cout = w.shape[0]
    if w.dim() != 2 or w.shape[1] != 9 * (c1 + c2) or not w.is_contiguous():
        raise ValueError(f"w must be contiguous [Cout, 9*Cin]={cout, 9 * (c1 + c2)}, got {tuple(w.shape)}")
    if stride == 2 and (h % 2 or wd % 2):
        raise ValueError("stride-2 conv needs even H and W")
    ho, wo = h // stride, wd // stride
    if out is None:
        out = torch.empty((n, ho, wo, cout), device=x.device, dtype=BF16)
    _req(out, "out")
    if tuple(out.shape) != (n, ho, wo, cout) or not out.is_contiguous():
        raise ValueError("out must be contiguous [N,Ho,Wo,Cout]")
    if residual is not None:
        _req(residual, "residual")
        if tuple(residual.shape) != tuple(out.shape) or not residual.is_contiguous():
            raise ValueError("residual must be contiguous and shaped like out")
    if bias is not None:
        _req(bias, "bias")
    if img_bias is not None:
        _req(img_bias, "img_bias", torch.float32)
        if tuple(img_bias.shape) != (n, cout) or img_bias.stride(1) != 1:
            raise ValueError("img_bias must be fp32 [N, Cout] with unit inner stride")
    if film is not None and film[0].shape[0] != n:  # camera FiLM of a block output, per image
        raise ValueError("film rows must match the image count")
    ex, _keep = _extras(None, None, film, cout, n, x.device)
    check(
        lib().mvd_conv3x3_ex_bf16(_p(x), c1, _p(x2), c2, _p(w), _p(bias), _p(img_bias),
                                  0 if img_bias is None else img_bias.stride(0), _p(residual), _p(out), n, ho, wo,
                                  cout, stride, tile_n, ctypes.byref(ex), _stream()),
        "mvd_conv3x3_ex_bf16",
    )
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None, split_tail: bool = True, co_units: int = 0) -> torch.Tensor:
    """q: [B,Sq,>=heads*64] view, k/v: [B,Skv,...] views (unit inner stride); head h = columns h*64.. .
    k/v may be batch-broadcast views (stride(0) == 0, e.g. `kv.expand(B, -1, -1)`): one K/V sequence shared by all
    batch entries (cross-view reference mode).
    split_tail: let the launch split the units of its last partial wave along S_kv (needs the per-stream workspace);
    co_units: 256-row units of another attention launch running concurrently on a different stream (see the header)."""
    for t, nme in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, nme)
        if t.dim() != 3 or t.stride(2) != 1 or t.shape[2] != heads * 64:
            raise ValueError(f"{nme} must be [B,S,{heads * 64}] with unit inner stride, got {tuple(t.shape)}")
    B, Sq, C = q.shape
    Skv = k.shape[1]
    if k.shape[0] != B or v.shape[0] != B or v.shape[1] != Skv:
        raise ValueError(f"q/k/v batch or length mismatch: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)}")
    if out is None:
        out = torch.empty((B, Sq, C), device=q.device, dtype=BF16)
    _req(out, "out")
    if tuple(out.shape) != (B, Sq, C) or out.stride(2) != 1:
        raise ValueError("out must be [B,Sq,heads*64] with unit inner stride")
    if scale is None:
        scale = 0.125
    ws_ptr, ws_bytes = None, 0
    if split_tail:
        ws = _workspace(q.device, _attn_ws_bytes() // 4, "attn")  # per (device, stream): never shared by the two branches
        ws_ptr, ws_bytes = ws.data_ptr(), ws.numel() * 4
    check(
        lib().mvd_attention_bf16_ws(q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0),
                                    v.data_ptr(), v.stride(1), v.stride(0), out.data_ptr(), out.stride(1),
                                    out.stride(0), B, heads, Sq, Skv, float(scale), ws_ptr, ws_bytes, int(co_units),
                                    _stream()),
        "mvd_attention_bf16_ws",
    )
    return out


_ATTN_WS_BYTES = None


ATTN_PERSIST_MAX_KV = 96 * 128  # csrc/attn.cu ATT_PERSIST_MAX_BLOCKS


def attention_is_persistent(batch: int, heads: int, s_q: int, s_kv: int, sms: int) -> bool:
    """Mirror of the launcher's choice (csrc/attn.cu): at least one wave of 256-row units, two-tile kernel, items of at
    most ATT_PERSIST_MAX_BLOCKS KV blocks -> one persistent CTA per SM that owns the machine for the launch."""
    return ATTN_PERSIST and s_q >= 512 and s_kv <= ATTN_PERSIST_MAX_KV and attention_units(batch, heads, s_q) >= sms


def attention_units(batch: int, heads: int, s_q: int) -> int:
    """Scheduling units (256-row tile pair, head, batch) of an attention launch — the `co_units` of its sibling."""
    return ((s_q + 255) // 256) * heads * batch


def _attn_ws_bytes() -> int:
    global _ATTN_WS_BYTES
    if _ATTN_WS_BYTES is None:
        _ATTN_WS_BYTES = int(lib().mvd_attention_workspace_bytes())
    return _ATTN_WS_BYTES


# ----------------------------------------------------------------------------------------------------------
# bandwidth-bound kernels
# ----------------------------------------------------------------------------------------------------------
F32 = torch.float32
_workspaces: dict = {}


def _workspace(device, nfloats: int, tag: str = "ws") -> torch.Tensor:
    """Persistent zero-initialised fp32 scratch per (device, stream, tag): stable addresses (CUDA-graph friendly) and
    never shared between streams that could run concurrently (GroupNorm keeps ticket counters in it)."""
    key = (str(device), _stream(), tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nfloats:
        ws = torch.zeros(max(nfloats, 1 << 16), device=device, dtype=F32)  # zero: GroupNorm ticket counters
        _workspaces[key] = ws
    return ws


def _contig(t: torch.Tensor, name: str, dtype=BF16):
    _req(t, name, dtype)
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def groupnorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 32, eps: float = 1e-5,
              silu: bool = False, x2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: NHWC [N,H,W,C1] or [N,HW,C1] (x2 concatenated along C); returns a dense tensor of x's rank."""
    _contig(x, "x")
    n, c1 = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * c1)
    c2 = 0
    if x2 is not None:
        _contig(x2, "x2")
        c2 = x2.shape[-1]
        if x2.shape[:-1] != x.shape[:-1]:
            raise ValueError("x2 must match x except for channels")
    _contig(gamma, "gamma")
    _contig(beta, "beta")
    if gamma.numel() != c1 + c2 or beta.numel() != c1 + c2:
        raise ValueError("gamma/beta size mismatch")
    if out is None:
        out = torch.empty(tuple(x.shape[:-1]) + (c1 + c2,), device=x.device, dtype=BF16)
    _contig(out, "out")
    need = lib().mvd_groupnorm_workspace_floats(n, hw, groups)
    ws = _workspace(x.device, need, "gn")
    check(
        lib().mvd_groupnorm_bf16(_p(x), c1, _p(x2), c2, _p(gamma), _p(beta), _p(out), n, hw, groups, float(eps),
                                 int(silu), _p(ws), ws.numel(), _stream()),
        "mvd_groupnorm_bf16",
    )
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [..., C] bf16 with dense rows; LayerNorm over C."""
    _req(x, "x")
    C = x.shape[-1]
    x2d = x.reshape(-1, C)
    px, ldx, M, _ = _rows2d(x2d, "x")
    if out is None:
        out = torch.empty_like(x2d)
    o2d = out.reshape(-1, C)
    po, ldo, _, _ = _rows2d(o2d, "out")
    check(lib().mvd_layernorm_bf16(px, ldx, _p(gamma), _p(beta), po, ldo, M, C, float(eps), _stream()),
          "mvd_layernorm_bf16")
    return out.view(x.shape)


def layernorm_f32(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                  silu: bool = False) -> torch.Tensor:
    _contig(x, "x", F32)
    M, C = x.shape
    out = torch.empty_like(x)
    check(lib().mvd_layernorm_f32(_p(x), _p(gamma), _p(beta), _p(out), M, C, float(eps), int(silu), _stream()),
          "mvd_layernorm_f32")
    return out


def refnorm(x: torch.Tensor, per_pixel: bool, out: Optional[torch.Tensor] = None, replication: int = 1) -> torch.Tensor:
    """x: bf16 [B,S,C] channels-last reference features -> normalised (attention.py:95-103 semantics).
    replication > 1 (3-D form only): x stands for that many identical copies along the batch; the statistics are those
    of the replicated tensor, ONE normalised copy is returned."""
    _contig(x, "x")
    B, S, C = x.shape
    if out is None:
        out = torch.empty_like(x)
    need = lib().mvd_refnorm_workspace_floats(C)
    ws = _workspace(x.device, need, "refnorm")
    check(lib().mvd_refnorm_replicated_bf16(_p(x), _p(out), B, S, C, int(per_pixel), int(replication), _p(ws),
                                            ws.numel(), _stream()), "mvd_refnorm_replicated_bf16")
    return out


def film(x: torch.Tensor, mod: torch.Tensor, strength: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: NHWC bf16 [N,...,C]; mod: fp32 [V, 2C]."""
    _contig(x, "x")
    _contig(mod, "mod", F32)
    n, C = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * C)
    V = mod.shape[0]
    if mod.shape[1] != 2 * C:
        raise ValueError("mod must be [V, 2C]")
    if n % V != 0:
        raise ValueError(f"batch {n} is not a multiple of the number of cameras {V}")
    if out is None:
        out = torch.empty_like(x)
    check(lib().mvd_film_bf16(_p(x), _p(out), _p(mod), n, V, hw, C, float(strength), _stream()), "mvd_film_bf16")
    return out


def small_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, silu_in: bool = False,
                 silu_out: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: fp32 [M<=16, K]; w: bf16 [N,K]; returns fp32 [M,N]."""
    _req(x, "x", F32)
    _contig(w, "w")
    px, ldx, M, K = _rows2d(x, "x")
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError("inner dimension mismatch")
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=F32)
    po, ldo, _, _ = _rows2d(out, "out")
    check(lib().mvd_small_linear_f32(px, ldx, _p(w), _p(bias), po, ldo, M, N, K, int(silu_in), int(silu_out),
                                     _stream()), "mvd_small_linear_f32")
    return out


def timestep_embedding(t: torch.Tensor, batch: int, dim: int = 320) -> torch.Tensor:
    _contig(t, "timesteps", F32)
    out = torch.empty((batch, dim), device=t.device, dtype=F32)
    check(lib().mvd_timestep_embedding_f32(_p(t), t.numel(), _p(out), batch, dim, _stream()),
          "mvd_timestep_embedding_f32")
    return out


def camera_front(src: torch.Tensor, tgt: torch.Tensor, pos_enc_dim: int, max_freq: float):
    _contig(src, "source_camera", F32)
    _contig(tgt, "target_camera", F32)
    V = src.shape[0]
    r = torch.empty((V, 9), device=src.device, dtype=F32)
    enc = torch.empty((V, 6 * pos_enc_dim), device=src.device, dtype=F32)
    t_rel = torch.empty((V, 3), device=src.device, dtype=F32)
    check(lib().mvd_camera_front_f32(_p(src), _p(tgt), _p(r), _p(enc), _p(t_rel), V, pos_enc_dim, float(max_freq),
                                     _stream()), "mvd_camera_front_f32")
    return r, enc, t_rel


def conv_in(latents: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, n_img: int, mod: Optional[torch.Tensor] = None,
            strength: float = 1.0) -> torch.Tensor:
    """latents: fp32 NCHW [n_lat,4,H,W]; w: bf16 [Cout,3,3,4]; returns NHWC bf16 [n_img,H,W,Cout]."""
    _contig(latents, "latents", F32)
    _contig(w, "w")
    n_lat, c, H, W = latents.shape
    if c != 4:
        raise ValueError("conv_in expects 4 latent channels")
    cout = w.shape[0]
    V = 0
    if mod is not None:
        _contig(mod, "mod", F32)
        V = mod.shape[0]
        if mod.shape[1] != 8 or n_img % V != 0:
            raise ValueError("mod must be [V, 8] with n_img a multiple of V")
    out = torch.empty((n_img, H, W, cout), device=latents.device, dtype=BF16)
    check(lib().mvd_conv_in_f32_bf16(_p(latents), n_lat, _p(mod), V, float(strength), _p(w), _p(bias), _p(out), n_img,
                                     H, W, cout, _stream()), "mvd_conv_in_f32_bf16")
    return out


def conv_out(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """x: NHWC bf16 [N,H,W,Cin]; w: bf16 [4,3,3,Cin]; returns fp32 NCHW [N,4,H,W]."""
    _contig(x, "x")
    _contig(w, "w")
    n, H, W, cin = x.shape
    out = torch.empty((n, 4, H, W), device=x.device, dtype=F32)
    check(lib().mvd_conv_out_bf16_f32(_p(x), _p(w), _p(bias), _p(out), n, H, W, cin, _stream()),
          "mvd_conv_out_bf16_f32")
    return out


def head4_to_nchw(x: torch.Tensor) -> torch.Tensor:
    """x: NHWC bf16 [N,H,W,C>=4] -> fp32 NCHW [N,4,H,W] of its first 4 channels (the tail of the 32-column conv_out)."""
    _contig(x, "x")
    n, H, W, C = x.shape
    out = torch.empty((n, 4, H, W), device=x.device, dtype=F32)
    check(lib().mvd_head4_to_nchw_f32(_p(x), C, _p(out), n, H * W, _stream()), "mvd_head4_to_nchw_f32")
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    _contig(x, "x")
    n, H, W, C = x.shape
    out = torch.empty((n, 2 * H, 2 * W, C), device=x.device, dtype=BF16)
    check(lib().mvd_upsample_nearest2x_bf16(_p(x), _p(out), n, H, W, C, _stream()), "mvd_upsample_nearest2x_bf16")
    return out


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _contig(a, "a")
    _contig(b, "b")
    if a.shape != b.shape:
        raise ValueError("shape mismatch")
    if out is None:
        out = torch.empty_like(a)
    check(lib().mvd_add_bf16(_p(a), _p(b), _p(out), a.numel(), _stream()), "mvd_add_bf16")
    return out


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _contig(x, "x", F32)
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    check(lib().mvd_cast_f32_bf16(_p(x), _p(out), x.numel(), _stream()), "mvd_cast_f32_bf16")
    return out


_DT = {torch.float32: 0, torch.bfloat16: 1}


def transpose_batched(x: torch.Tensor, out_dtype=BF16) -> torch.Tensor:
    """x: contiguous [B, R, C] (fp32 or bf16) -> [B, C, R] of out_dtype. NCHW<->NHWC with R/C = channels/pixels."""
    if x.dtype not in _DT or out_dtype not in _DT:
        raise ValueError("transpose supports fp32 and bf16")
    _contig(x, "x", x.dtype)
    B, R, C = x.shape
    out = torch.empty((B, C, R), device=x.device, dtype=out_dtype)
    check(lib().mvd_transpose_batched(_p(x), _p(out), B, R, C, _DT[x.dtype], _DT[out_dtype], _stream()),
          "mvd_transpose_batched")
    return out


def cfg_ddpm_step(model_out: torch.Tensor, latents: torch.Tensor, noise: Optional[torch.Tensor], cfg: int,
                  guidance: float, sqrt_abar: float, sqrt_1m_abar: float, c_x0: float, c_xt: float,
                  sigma: float) -> torch.Tensor:
    """In-place DDPM step on fp32 latents; model_out fp32 [cfg * latents.numel()]."""
    _contig(model_out, "model_out", F32)
    _contig(latents, "latents", F32)
    n = latents.numel()
    if model_out.numel() != cfg * n:
        raise ValueError("model_out must hold cfg * latents elements")
    if noise is not None:
        _contig(noise, "noise", F32)
    check(lib().mvd_cfg_ddpm_step_f32(_p(model_out), _p(latents), _p(noise), n, cfg, float(guidance), float(sqrt_abar),
                                      float(sqrt_1m_abar), float(c_x0), float(c_xt), float(sigma), _stream()),
          "mvd_cfg_ddpm_step_f32")
    return latents


def cfg_ddpm_step_table(model_out: torch.Tensor, latents: torch.Tensor, noise_table: Optional[torch.Tensor], cfg: int,
                        guidance: float, coef_table: torch.Tensor, step_idx: torch.Tensor) -> torch.Tensor:
    """Graph-replayable step: scalars come from coef_table[step_idx] on the device (see include/mvd_b200.h)."""
    _contig(model_out, "model_out", F32)
    _contig(latents, "latents", F32)
    _contig(coef_table, "coef_table", F32)
    _contig(step_idx, "step_idx", torch.int32)
    n = latents.numel()
    if model_out.numel() != cfg * n or coef_table.dim() != 2 or coef_table.shape[1] != 8:
        raise ValueError("bad shapes for the table-driven step")
    if noise_table is not None:
        _contig(noise_table, "noise_table", F32)
        if noise_table.numel() != coef_table.shape[0] * n:
            raise ValueError("noise_table must be [steps, latents.numel()]")
    check(lib().mvd_cfg_ddpm_step_table_f32(_p(model_out), _p(latents), _p(noise_table), n, cfg, float(guidance),
                                            _p(coef_table), _p(step_idx), _stream()), "mvd_cfg_ddpm_step_table_f32")
    return latents


def advance_step(step_idx: torch.Tensor, coef_table: torch.Tensor, timestep_out: torch.Tensor) -> None:
    _contig(step_idx, "step_idx", torch.int32)
    _contig(coef_table, "coef_table", F32)
    _contig(timestep_out, "timestep_out", F32)
    check(lib().mvd_advance_step(_p(step_idx), _p(coef_table), _p(timestep_out), coef_table.shape[0], _stream()),
          "mvd_advance_step")


def advance_step_rows(step_idx: torch.Tensor, coef_table: torch.Tensor, timestep_out: torch.Tensor,
                      row_table: torch.Tensor, row_out: torch.Tensor) -> None:
    """advance_step + copy of the new step's row of a per-schedule fp32 table [steps, row_len] into row_out."""
    _contig(step_idx, "step_idx", torch.int32)
    _contig(coef_table, "coef_table", F32)
    _contig(timestep_out, "timestep_out", F32)
    _contig(row_table, "row_table", F32)
    _contig(row_out, "row_out", F32)
    if row_table.dim() != 2 or row_table.shape[0] != coef_table.shape[0] or row_out.numel() != row_table.shape[1]:
        raise ValueError("row_table must be [steps, row_len] and row_out hold row_len floats")
    check(lib().mvd_advance_step_rows(_p(step_idx), _p(coef_table), _p(timestep_out), coef_table.shape[0],
                                      _p(row_table), _p(row_out), row_table.shape[1], _stream()),
          "mvd_advance_step_rows")


def set_launch_overlap(on: bool) -> None:
    """Programmatic dependent launch for all kernels of the library (process-wide; see the header)."""
    check(lib().mvd_set_launch_overlap(int(bool(on))), "mvd_set_launch_overlap")


def kernel_launch_count() -> int:
    return int(lib().mvd_kernel_launch_count())
