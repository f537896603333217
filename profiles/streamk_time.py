"""Stream-K on / off for every non-GEGLU GEMM / conv shape class of the step at B samples per GPU (B = 1, 2, 8):
kernel time from a CUDA graph of 10 calls, min of 5 replays. Columns: whole tiles only (workspace withheld), stream-K at
the cost model's tile width, stream-K at forced tile widths.  usage: python profiles/streamk_time.py [B ...]"""
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
COLS = (("whole", False, 0), ("sk", True, 0), ("sk64", True, 64), ("sk128", True, 128), ("sk160", True, 160), ("sk256", True, 256))
def row(name, make):
    cells = []
    for _, sk, bn in COLS:
        ops.SPLIT_K = sk
        try:
            cells.append(f"{timeit(make(bn)) * 1e3:8.1f}")
        except Exception:
            cells.append(f"{'-':>8}")
    ops.SPLIT_K = True
    print(f"{name:44s} " + " ".join(cells), flush=True)
total = {}
for B in [int(a) for a in sys.argv[1:]] or [1, 2, 8]:
    print(f"--- {B} sample(s) per GPU\n{'shape':44s} " + " ".join(f"{c[0]:>8}" for c in COLS) + "   (us)")
    for lvl, (hw, C) in enumerate([(4096, 320), (1024, 640), (256, 1280), (64, 1280)]):
        M = B * hw
        for (n, k, res, name) in [(4 * C, C, False, "attn1 q,k,v,q_ref"), (C, 2 * C, True, "out-proj K=2C"), (C, 4 * C, True, "ff2"),
                                  (C, C, False, "proj_in/out")]:
            a, w, b = rnd(M, k), rnd(n, k, scale=k ** -0.5), rnd(n)
            r = rnd(M, n) if res else None
            row(f"linear {name:18s} M={M} N={n} K={k}",
                lambda bn, a=a, w=w, b=b, r=r: (lambda: ops.linear(a, w, bias=b, residual=r, tile_n=bn)))
    for (h, c1, c2) in [(64, 320, 320), (64, 640, 320), (64, 960, 320), (32, 320, 640), (32, 640, 640), (32, 1280, 640), (32, 1920, 640),
                        (16, 640, 1280), (16, 1280, 1280), (16, 2560, 1280), (16, 1920, 1280), (8, 1280, 1280), (8, 2560, 1280)]:
        x, w = rnd(B, h, h, c1), rnd(c2, 9 * c1, scale=(9 * c1) ** -0.5)
        b, r = rnd(c2), rnd(B, h, h, c2)
        ib = torch.randn(B, c2, device=dev)
        row(f"conv3x3 {B}x{h}x{h} {c1}->{c2}",
            lambda bn, x=x, w=w, b=b, r=r, ib=ib: (lambda: ops.conv3x3(x, w, bias=b, img_bias=ib, residual=r, tile_n=bn)))
