"""Algorithmic FLOP counter (2*MAC) for one UNet(+adapter) forward per sample — SURVEY.md Appendix C.
Walks the ORACLE's module tree with symbolic shapes (no tensors are allocated)."""
import torch
import torch.nn as nn


def unet_flops(L: int, skv_factor: int = 1, cfg=None):
    from oracle.sd21_unet import (Attention, DownBlock, MidBlock, ResnetBlock2D, Transformer2DModel, UNet2DConditionModel,
                                  UpBlock)

    with torch.device("meta"):
        u = UNet2DConditionModel(**(cfg or {}))
    f = dict(conv=0, linear=0, self_attn=0, text_attn=0, ref_proj=0, ref_attn=0)

    def conv(m: nn.Conv2d, hw_out):
        k = m.kernel_size[0] * m.kernel_size[1]
        f["conv"] += 2 * m.in_channels * m.out_channels * k * hw_out

    def lin(m: nn.Linear, rows, key="linear"):
        f[key] += 2 * m.in_features * m.out_features * rows

    def resnet(r: ResnetBlock2D, hw):
        conv(r.conv1, hw)
        conv(r.conv2, hw)
        lin(r.time_emb_proj, 1)
        if r.conv_shortcut is not None:
            conv(r.conv_shortcut, hw)

    def transformer(t: Transformer2DModel, hw):
        lin(t.proj_in, hw)
        lin(t.proj_out, hw)
        for b in t.transformer_blocks:
            c = b.attn1.to_q.in_features
            for a, s_kv, key in ((b.attn1, hw, "self_attn"), (b.attn2, 77, "text_attn")):
                lin(a.to_q, hw)
                lin(a.to_k, s_kv)
                lin(a.to_v, s_kv)
                lin(a.to_out[0], hw)
                f[key] += 4 * hw * s_kv * c
                # adapter processor on this Attention (reference attention.py:125-158)
                s_ref = hw * skv_factor
                f["ref_proj"] += 2 * c * c * (hw + s_ref + s_ref + hw)
                f["ref_attn"] += 4 * hw * s_ref * c
            lin(b.ff.net[0].proj, hw)
            lin(b.ff.net[2], hw)

    side = L
    conv(u.conv_in, side * side)
    lin(u.time_embedding.linear_1, 1)
    lin(u.time_embedding.linear_2, 1)
    for blk in u.down_blocks:
        hw = side * side
        for i, r in enumerate(blk.resnets):
            resnet(r, hw)
            if blk.has_attn:
                transformer(blk.attentions[i], hw)
        if blk.downsamplers is not None:
            side //= 2
            conv(blk.downsamplers[0].conv, side * side)
    hw = side * side
    resnet(u.mid_block.resnets[0], hw)
    transformer(u.mid_block.attentions[0], hw)
    resnet(u.mid_block.resnets[1], hw)
    for blk in u.up_blocks:
        hw = side * side
        for i, r in enumerate(blk.resnets):
            resnet(r, hw)
            if blk.has_attn:
                transformer(blk.attentions[i], hw)
        if blk.upsamplers is not None:
            side *= 2
            conv(blk.upsamplers[0].conv, side * side)
    conv(u.conv_out, side * side)
    f["base"] = f["conv"] + f["linear"] + f["self_attn"] + f["text_attn"]
    f["adapter"] = f["ref_proj"] + f["ref_attn"]
    f["total"] = f["base"] + f["adapter"]
    return f
