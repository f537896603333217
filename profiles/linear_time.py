import os, sys
sys.path.insert(0, "/root/repo")
import torch
from mvd_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, device="cuda", generator=g) * scale).to(torch.bfloat16)
def timeit(fns, reps=3):
    for f in fns: f()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            for f in fns: f()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / (reps * len(fns)))
    return min(ts)
print("MVD_GEMM_STORE_HINT =", os.environ.get("MVD_GEMM_STORE_HINT", "0"))
M = 32768
for (n, k, res, name) in [(1280, 320, False, "qkv+q_ref"), (640, 320, False, "attn2 q"), (320, 640, True, "out-proj"), (320, 1280, True, "ff2"), (320, 320, False, "proj")]:
    # 4 independent operand sets used round-robin (working set > L2), as in the step where every launch sees new data
    sets = [(rnd(M, k), rnd(n, k, scale=k ** -0.5), rnd(n), rnd(M, n) if res else None) for _ in range(4)]
    fns = [(lambda a=a, w=w, b=b, r=r: ops.linear(a, w, bias=b, residual=r)) for (a, w, b, r) in sets]
    ms = timeit(fns)
    print(f"  {name:10s} M={M} N={n} K={k}: {ms*1e3:6.1f} us  {2.0*M*n*k/ms/1e9:5.0f} TFLOP/s", flush=True)
