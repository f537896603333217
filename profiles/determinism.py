import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def rnd(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)
def check(name, fn, n=6):
    ref = fn().clone()
    bad = 0
    for _ in range(n):
        junk = torch.randn(1 << 22, device=dev)  # perturb allocator / caches
        out = fn()
        bad += int(not torch.equal(out, ref))
        del junk
    print(f"{name:40s} nondeterministic runs: {bad}/{n}")
for (n, hw, c) in [(2, 256, 64), (2, 64, 128), (2, 4, 128), (8, 4096, 320)]:
    x = rnd(n, hw, c); gm, bt = rnd(c), rnd(c)
    check(f"groupnorm {n}x{hw}x{c}", lambda: ops.groupnorm(x, gm, bt, silu=True))
for (M, N, K) in [(512, 64, 64), (8, 128, 128), (2048, 320, 320), (32768, 320, 640)]:
    a, w, b, r = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N), rnd(M, N)
    check(f"linear {M}x{N}x{K} bias+res", lambda: ops.linear(a, w, bias=b, residual=r))
    check(f"linear {M}x{N}x{K} plain", lambda: ops.linear(a, w))
for (n, h, c1, c2) in [(2, 16, 64, 64), (2, 2, 128, 128), (8, 64, 320, 320)]:
    x, w = rnd(n, h, h, c1), rnd(c2, 9 * c1, scale=(9 * c1) ** -0.5)
    b, r, ib = rnd(c2), rnd(n, h, h, c2), torch.randn(n, c2, device=dev)
    check(f"conv {n}x{h}x{h} {c1}->{c2}", lambda: ops.conv3x3(x, w, bias=b, img_bias=ib, residual=r))
for (B, H, Sq, Skv) in [(2, 1, 256, 256), (2, 2, 64, 77), (2, 2, 4, 4), (8, 5, 4096, 4096), (8, 10, 1024, 1024)]:
    q, k, v = rnd(B, Sq, H * 64), rnd(B, Skv, H * 64), rnd(B, Skv, H * 64)
    check(f"attention B{B} h{H} {Sq}x{Skv}", lambda: ops.attention(q, k, v, H))
x = rnd(512, 128); gm, bt = rnd(128), rnd(128)
check("layernorm", lambda: ops.layernorm(x, gm, bt))
