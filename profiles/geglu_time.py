import os, sys
sys.path.insert(0, "/root/repo")
import torch
from mvd_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, device="cuda", generator=g) * scale).to(torch.bfloat16)
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
for B in (1, 2, 8):
    for (hw, C) in [(4096, 320), (1024, 640), (256, 1280), (64, 1280)]:
        M, K, N = B * hw, C, 8 * C
        a, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), rnd(N)
        cells = []
        for tn in (256, 128, 64):
            try:
                ms = timeit(lambda: ops.linear(a, w, bias=b, geglu=True, tile_n=tn))
                cells.append(f"{ms*1e3:7.1f}")
            except Exception as e:
                cells.append("      -")
        # plain linear of the same size for comparison (no GELU, N outputs)
        ms = timeit(lambda: ops.linear(a, w, bias=b))
        print(f"B={B} geglu M={M} N={N} K={K}: bn256/128/64 = {' '.join(cells)} us   plain linear {ms*1e3:6.1f} us  ({2.0*M*N*K/ms/1e9:5.0f} TFLOP/s)", flush=True)
