"""Bitwise determinism of the kernels that exchange partial results through workspaces (stream-K GEMM / conv, KV-split
and persistent attention, cluster / window GroupNorm), repeated back to back and concurrently on two streams (each
stream has its own workspaces): every repetition must reproduce the first result exactly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, device="cuda", generator=g) * scale).to(torch.bfloat16)
cases = {}
x, w = rnd(1, 32, 32, 1280), rnd(640, 9 * 1280, scale=(9 * 1280) ** -0.5)
cases["conv stream-K 1x32x32 1280->640"] = lambda: ops.conv3x3(x, w)
x2, w2 = rnd(8, 64, 64, 320), rnd(320, 9 * 320, scale=(9 * 320) ** -0.5)
cases["conv hybrid tail 8x64x64 320->320"] = lambda: ops.conv3x3(x2, w2)
a, wl = rnd(256, 5120), rnd(1280, 5120, scale=5120 ** -0.5)
cases["linear stream-K 256x5120x1280"] = lambda: ops.linear(a, wl)
q, kv = rnd(8, 4096, 320), rnd(8, 4096, 640)
cases["attention persistent B=8"] = lambda: ops.attention(q, kv[:, :, :320], kv[:, :, 320:], 5)
q1, kv1 = rnd(1, 4096, 320), rnd(1, 4096, 640)
cases["attention KV-split B=1"] = lambda: ops.attention(q1, kv1[:, :, :320], kv1[:, :, 320:], 5)
xg = rnd(8, 4096, 320); gm, bt = rnd(320, scale=0.2) + 1, rnd(320, scale=0.2)
cases["groupnorm window 8x4096x320"] = lambda: ops.groupnorm(xg, gm, bt, groups=32, eps=1e-5, silu=True)
xg1 = rnd(1, 4096, 320)
cases["groupnorm cluster 1x4096x320"] = lambda: ops.groupnorm(xg1, gm, bt, groups=32, eps=1e-5, silu=True)
side = torch.cuda.Stream()
bad = 0
for name, fn in cases.items():
    ref = fn().clone()
    torch.cuda.synchronize()
    diffs = 0
    for rep in range(30):
        outs = []
        with torch.cuda.stream(side):
            side.wait_stream(torch.cuda.current_stream())
            o_side = fn()
        o_main = fn()  # runs concurrently with the side-stream launch
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        diffs += int(not torch.equal(o_main, ref)) + int(not torch.equal(o_side, ref))
    print(f"{name:40s} {60 - diffs}/60 bit-identical", flush=True)
    bad += diffs
print("stress_determinism:", "ok" if bad == 0 else f"{bad} MISMATCHES")
sys.exit(1 if bad else 0)
