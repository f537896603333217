"""Stand-alone launches of the hot kernels at their configs[1] shapes, for ncu (`--set full -k regex:...`).
usage: python profiles/prof_kernels.py attn|gemm|conv [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops

what = sys.argv[1] if len(sys.argv) > 1 else "attn"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)


if what == "attn":
    B, H, S = 8, 5, 4096
    C = H * 64
    qkv = rnd(B, S, 3 * C)
    out = torch.empty(B, S, C, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, out=out)
elif what == "gemm":
    M = 8 * 4096
    for (n, k, kw) in [(1280, 320, {}), (320, 640, {}), (2560, 320, dict(geglu=True, tile_n=256)), (320, 1280, {})]:
        a, w = rnd(M, k), rnd(n, k, scale=k ** -0.5)
        b = rnd(n)
        r = None if kw else rnd(M, n)
        for _ in range(reps):
            ops.linear(a, w, bias=b, residual=r, **kw)
elif what == "conv":
    for (n, h, c1, c2) in [(8, 64, 320, 320), (8, 32, 640, 640), (8, 16, 1280, 1280), (8, 8, 1280, 1280)]:
        x, w = rnd(n, h, h, c1), rnd(c2, 9 * c1, scale=(9 * c1) ** -0.5)
        b, r = rnd(c2), rnd(n, h, h, c2)
        ib = torch.randn(n, c2, device=dev)
        for _ in range(reps):
            ops.conv3x3(x, w, bias=b, img_bias=ib, residual=r)
torch.cuda.synchronize()
print("done", what)
