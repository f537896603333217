"""Thin torch-tensor wrappers over the C ABI (include/mvd_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream. Every function validates
device/dtype/layout, passes raw pointers to libmvd_b200.so and raises on any error. No op has a PyTorch
or CPU fallback.
"""
from __future__ import annotations

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
) -> torch.Tensor:
    """out[M,N] = [a|a2] @ w^T (+bias) (+row_group_bias[row // rows_per_group]) (+residual); optional GEGLU."""
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
    check(
        lib().mvd_linear_bf16(pa, lda, k1, pa2, lda2, k2, pw, ldw, _p(bias), pg, ldg, rows_per_group, pr, ldr, po, ldo,
                              M, N, int(geglu), tile_n, _stream()),
        "mvd_linear_bf16",
    )
    return out


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
        if tuple(img_bias.shape) != (n, cout) or not img_bias.is_contiguous():
            raise ValueError("img_bias must be contiguous fp32 [N, Cout]")
    check(
        lib().mvd_conv3x3_bf16(_p(x), c1, _p(x2), c2, _p(w), _p(bias), _p(img_bias), _p(residual), _p(out), n, ho, wo,
                               cout, stride, tile_n, _stream()),
        "mvd_conv3x3_bf16",
    )
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q: [B,Sq,>=heads*64] view, k/v: [B,Skv,...] views (unit inner stride); head h = columns h*64.. ."""
    for t, nme in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, nme)
        if t.dim() != 3 or t.stride(2) != 1 or t.shape[2] != heads * 64:
            raise ValueError(f"{nme} must be [B,S,{heads * 64}] with unit inner stride, got {tuple(t.shape)}")
    B, Sq, C = q.shape
    Skv = k.shape[1]
    if k.shape[0] != B or v.shape[0] != B or v.shape[1] != Skv:
        raise ValueError("q/k/v batch or length mismatch")
    if out is None:
        out = torch.empty((B, Sq, C), device=q.device, dtype=BF16)
    _req(out, "out")
    if tuple(out.shape) != (B, Sq, C) or out.stride(2) != 1:
        raise ValueError("out must be [B,Sq,heads*64] with unit inner stride")
    if scale is None:
        scale = 0.125
    check(
        lib().mvd_attention_bf16(q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0),
                                 v.data_ptr(), v.stride(1), v.stride(0), out.data_ptr(), out.stride(1), out.stride(0),
                                 B, heads, Sq, Skv, float(scale), _stream()),
        "mvd_attention_bf16",
    )
    return out
