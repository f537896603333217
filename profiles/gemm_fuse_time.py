"""Cost of the fused GEMM epilogue options at the step's linear shapes: plain, + row statistics of the output
(stats_out), + LayerNorm fold (ln), against the LayerNorm kernel they replace.  usage: python profiles/gemm_fuse_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
from mvd_b200.unet import fold_layernorm

g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (M, C) in [(32768, 320), (8192, 640), (2048, 1280), (512, 1280)]:
    x, r = rnd(M, C), rnd(M, C)
    norm = torch.nn.LayerNorm(C).cuda()
    gam, bet = norm.weight.detach().to(torch.bfloat16), norm.bias.detach().to(torch.bfloat16)
    wo, bo = rnd(C, C, scale=C ** -0.5), rnd(C)
    _, st = ops.linear(x, wo, bias=bo, residual=r, want_stats=True)
    print(f"M={M} C={C}")
    print(f"  layernorm kernel                      {timeit(lambda: ops.layernorm(x, gam, bet, 1e-5)):7.1f} us")
    print(f"  out-proj C->C +bias +res              {timeit(lambda: ops.linear(x, wo, bias=bo, residual=r)):7.1f} us")
    print(f"  out-proj C->C +bias +res +stats_out   {timeit(lambda: ops.linear(x, wo, bias=bo, residual=r, want_stats=True)):7.1f} us")
    for N, name in ((4 * C, "qkv+q_ref 4C"), (2 * C, "q+q_ref 2C")):
        w = rnd(N, C, scale=C ** -0.5)
        wg, cs, cst = fold_layernorm(w, norm)
        print(f"  {name:14s} plain                  {timeit(lambda: ops.linear(x, w)):7.1f} us")
        print(f"  {name:14s} + LayerNorm fold       {timeit(lambda: ops.linear(x, wg, row_group_bias=cst, rows_per_group=M, ln=ops.LNFold(st, cs, 1e-5))):7.1f} us")
    w1, b1 = rnd(8 * C, C, scale=C ** -0.5), rnd(8 * C)
    wg, cs, cst = fold_layernorm(w1, norm, b1)
    b1f = cst.view(-1).to(torch.bfloat16)
    print(f"  geglu 8C       plain                  {timeit(lambda: ops.linear(x, w1, bias=b1, geglu=True, tile_n=256)):7.1f} us")
    print(f"  geglu 8C       + LayerNorm fold       {timeit(lambda: ops.linear(x, wg, bias=b1f, geglu=True, tile_n=256, ln=ops.LNFold(st, cs, 1e-5))):7.1f} us")
