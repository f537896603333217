"""Per-kernel cost inside a replayed CUDA graph for the small launches of a view-sharded rank (1 sample per GPU): chains
of 100 dependent launches of one op, and of alternating ops with different shared-memory footprints.
usage: python profiles/chain_latency.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def graph_time(fn, n=100, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / n * 1e3


M, C = 4096, 320
x, w, b, r = rnd(M, C), rnd(C, C, scale=C ** -0.5), rnd(C), rnd(M, C)
x4 = rnd(1, 64, 64, C)
gam, bet = rnd(C), rnd(C)
w9 = rnd(C, 9 * C, scale=(9 * C) ** -0.5)
buf = [x]


def lin():
    buf[0] = ops.linear(buf[0], w, bias=b, residual=r)


def gn():
    ops.groupnorm(x4, gam, bet, 32, 1e-5, silu=True)


def conv():
    ops.conv3x3(x4, w9, bias=b)


def add():
    ops.add(x, r)


x8 = rnd(1, 8, 8, 1280)
w8 = rnd(1280, 9 * 1280, scale=(9 * 1280) ** -0.5)


def conv8():
    ops.conv3x3(x8, w8)


def lin_gn():
    lin()
    gn()


for name, fn, per in [("add (elementwise, no smem)", add, 1), ("linear 4096x320x320 +bias +res", lin, 1),
                      ("groupnorm 1x64x64x320", gn, 1), ("conv3x3 1x64x64 320->320", conv, 1),
                      ("conv3x3 1x8x8 1280->1280 (split-K)", conv8, 1), ("linear then groupnorm (pair)", lin_gn, 2)]:
    for pdl in (0, 1):
        ops.set_launch_overlap(bool(pdl))
        print(f"{name:40s} PDL={pdl}: {graph_time(fn) / 1:7.2f} us per {'pair' if per == 2 else 'launch'}", flush=True)
