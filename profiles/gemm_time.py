"""Times the GEMM / conv kernel at the configs[1] shapes (CUDA events, min of 10; reports TFLOP/s)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)
def timeit(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
tot_f = tot_t = 0
M = 8 * 4096
for (n, k, kw, name) in [(1280, 320, {}, "qkv+qref 320"), (320, 640, {}, "out-proj K=2C 320"), (2560, 320, dict(geglu=True, tile_n=256), "geglu 320"),
                         (320, 1280, {}, "ff2 320"), (320, 320, {}, "proj 320")]:
    a, w, b = rnd(M, k), rnd(n, k, scale=k ** -0.5), rnd(n)
    r = None if kw else rnd(M, n)
    ms = timeit(lambda: ops.linear(a, w, bias=b, residual=r, **kw))
    fl = 2.0 * M * n * k
    print(f"linear {name:22s} M={M} N={n} K={k}: {ms*1e3:7.1f} us {fl/ms/1e9:7.0f} TFLOP/s")
M = 8 * 1024
for (n, k, kw, name) in [(2560, 640, {}, "qkv+qref 640"), (5120, 640, dict(geglu=True, tile_n=256), "geglu 640"), (640, 2560, {}, "ff2 640")]:
    a, w, b = rnd(M, k), rnd(n, k, scale=k ** -0.5), rnd(n)
    r = None if kw else rnd(M, n)
    ms = timeit(lambda: ops.linear(a, w, bias=b, residual=r, **kw))
    fl = 2.0 * M * n * k
    print(f"linear {name:22s} M={M} N={n} K={k}: {ms*1e3:7.1f} us {fl/ms/1e9:7.0f} TFLOP/s")
for (n, h, c1, c2) in [(8, 64, 320, 320), (8, 64, 640, 320), (8, 32, 640, 640), (8, 32, 1280, 640), (8, 16, 1280, 1280), (8, 16, 2560, 1280), (8, 8, 1280, 1280)]:
    x, w = rnd(n, h, h, c1), rnd(c2, 9 * c1, scale=(9 * c1) ** -0.5)
    b, r = rnd(c2), rnd(n, h, h, c2)
    ib = torch.randn(n, c2, device=dev)
    ms = timeit(lambda: ops.conv3x3(x, w, bias=b, img_bias=ib, residual=r))
    fl = 2.0 * n * h * h * 9 * c1 * c2
    print(f"conv3x3 {n}x{h}x{h} {c1}->{c2}: {ms*1e3:7.1f} us {fl/ms/1e9:7.0f} TFLOP/s")
