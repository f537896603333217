"""Times the bandwidth-bound kernels at configs[1] shapes; reports achieved GB/s of algorithmic bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def rnd(*shape, dtype=torch.bfloat16):
    return torch.randn(*shape, device=dev, generator=g).to(dtype)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
def timeit(fn, cold=False, reps=20):
    """Kernel time without host launch overhead: `reps` calls captured into one CUDA graph."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps): fn()
    ts = []
    for _ in range(5):
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps)
    return min(ts)
for (n, hw, c1, c2) in [(8, 4096, 320, 0), (8, 4096, 320, 320), (8, 1024, 640, 0), (8, 1024, 640, 640), (8, 256, 1280, 0), (8, 256, 1280, 1280), (8, 64, 1280, 1280)]:
    x1 = rnd(n, hw, c1); x2 = rnd(n, hw, c2) if c2 else None
    C = c1 + c2
    gm, bt = rnd(C), rnd(C)
    for cold in (False, True):
        ms = timeit(lambda: ops.groupnorm(x1, gm, bt, silu=True, x2=x2), cold)
        byts = 3.0 * n * hw * C * 2
        print(f"groupnorm {n}x{hw}x{c1}+{c2} {'cold' if cold else 'warm'}: {ms*1e3:7.1f} us  {byts/ms/1e6:7.0f} GB/s (2 launches)")
x = rnd(8 * 4096, 320); gm, bt = rnd(320), rnd(320)
ms = timeit(lambda: ops.layernorm(x, gm, bt)); print(f"layernorm 32768x320: {ms*1e3:.1f} us {2*x.numel()*2/ms/1e6:.0f} GB/s")
x = rnd(8, 4096, 320); mod = torch.randn(4, 640, device=dev)
ms = timeit(lambda: ops.film(x, mod, 1.0)); print(f"film 8x4096x320: {ms*1e3:.1f} us {2*x.numel()*2/ms/1e6:.0f} GB/s")
lat = torch.randn(4, 4, 64, 64, device=dev); w = rnd(320, 3, 3, 4); b = rnd(320)
ms = timeit(lambda: ops.conv_in(lat, w, b, n_img=8)); print(f"conv_in: {ms*1e3:.1f} us")
x = rnd(8, 64, 64, 320); w = rnd(4, 3, 3, 320); b = rnd(4)
ms = timeit(lambda: ops.conv_out(x, w, b)); print(f"conv_out: {ms*1e3:.1f} us")
xx = torch.randn(8, 1280, device=dev); w = rnd(20160, 1280); b = rnd(20160)
ms = timeit(lambda: ops.small_linear(xx, w, b, silu_in=True)); print(f"small_linear 8x20160x1280: {ms*1e3:.1f} us {w.numel()*2/ms/1e6:.0f} GB/s")
a = rnd(1024, 64); b = rnd(1024, 64); o = torch.empty_like(a)
ms = timeit(lambda: ops.add(a, b, out=o)); print(f"add tiny (launch floor): {ms*1e3:.1f} us")
x = rnd(8, 64, 2560); gm, bt = rnd(2560), rnd(2560)
ms = timeit(lambda: ops.groupnorm(x, gm, bt, silu=True)); print(f"groupnorm 8x64x2560 single-source: {ms*1e3:.1f} us")
x = rnd(8, 64, 64); gm, bt = rnd(64), rnd(64)
ms = timeit(lambda: ops.groupnorm(x, gm, bt, silu=True)); print(f"groupnorm 8x64x64: {ms*1e3:.1f} us")
