"""Sweeps the GEMM/conv tile width BN over every GEMM shape class of the configs[1] step (kernel time from a CUDA
graph of 10 calls, min of 5 replays). Output: one row per shape with us per BN and the BN the cost model picks
(tile_n=0). Used to calibrate pick_bn() in csrc/gemm.cu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)
def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps): fn()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps)
    return min(ts)
BNS = (0, 64, 128, 160, 256)
print(f"{'shape':44s} " + " ".join(f"{('auto' if b == 0 else b):>8}" for b in BNS) + "   (us; TFLOP/s of best)")
def row(name, fl, make):
    cells, best = [], 1e9
    for bn in BNS:
        try:
            ms = timeit(make(bn)); best = min(best, ms); cells.append(f"{ms*1e3:8.1f}")
        except Exception as e:  # unsupported width for this N
            cells.append(f"{'-':>8}")
    print(f"{name:44s} " + " ".join(cells) + f"   {fl/best/1e9:6.0f}")
for lvl, (hw, C) in enumerate([(4096, 320), (1024, 640), (256, 1280), (64, 1280)]):
    M = 8 * hw
    for (n, k, res, name) in [(4 * C, C, False, "attn1 q,k,v,q_ref"), (2 * C, C, False, "attn2 q,q_ref"), (C, 2 * C, True, "out-proj K=2C"),
                              (C, 4 * C, True, "ff2"), (C, C, False, "proj_in/out")]:
        a, w, b = rnd(M, k), rnd(n, k, scale=k ** -0.5), rnd(n)
        r = rnd(M, n) if res else None
        row(f"linear {name:18s} M={M} N={n} K={k}", 2.0 * M * n * k,
            lambda bn, a=a, w=w, b=b, r=r: (lambda: ops.linear(a, w, bias=b, residual=r, tile_n=bn)))
    n, k = 8 * C, C
    a, w, b = rnd(M, k), rnd(n, k, scale=k ** -0.5), rnd(n)
    def mk(bn, a=a, w=w, b=b):
        if bn == 160: raise ValueError
        return lambda: ops.linear(a, w, bias=b, geglu=True, tile_n=(bn or 256))
    row(f"linear {'ff1 geglu':18s} M={M} N={n} K={k}", 2.0 * M * n * k, mk)
for (n, h, c1, c2) in [(8, 64, 320, 320), (8, 64, 640, 320), (8, 64, 960, 320), (8, 32, 320, 640), (8, 32, 640, 640), (8, 32, 1280, 640), (8, 32, 1920, 640),
                       (8, 16, 640, 1280), (8, 16, 1280, 1280), (8, 16, 2560, 1280), (8, 16, 1920, 1280), (8, 8, 1280, 1280), (8, 8, 2560, 1280)]:
    x, w = rnd(n, h, h, c1), rnd(c2, 9 * c1, scale=(9 * c1) ** -0.5)
    b, r = rnd(c2), rnd(n, h, h, c2)
    ib = torch.randn(n, c2, device=dev)
    row(f"conv3x3 {n}x{h}x{h} {c1}->{c2}", 2.0 * n * h * h * 9 * c1 * c2,
        lambda bn, x=x, w=w, b=b, r=r, ib=ib: (lambda: ops.conv3x3(x, w, bias=b, img_bias=ib, residual=r, tile_n=bn)))
