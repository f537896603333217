"""GroupNorm(+SiLU) kernel time for the sites of the configs[1] step at 8 samples (CUDA graph of 10 calls, min of 5
replays), with the achieved fraction of the HBM copy peak for read-once + write-once traffic."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbps"]
except Exception:
    peak = 6550.7
print(f"MVD_GN_ROWS={os.environ.get('MVD_GN_ROWS', 'default')}  (HBM peak {peak:.0f} GB/s)")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for (n, hw, c1, c2) in [(8, 4096, 320, 0), (8, 4096, 320, 320), (8, 4096, 640, 320), (8, 1024, 640, 0), (8, 1024, 640, 640),
                        (8, 1024, 1280, 640), (8, 256, 1280, 0), (8, 256, 1280, 1280), (8, 64, 1280, 1280)]:
    x1 = torch.randn(n, hw, c1, device="cuda", generator=g).to(torch.bfloat16)
    x2 = torch.randn(n, hw, c2, device="cuda", generator=g).to(torch.bfloat16) if c2 else None
    C = c1 + c2
    gm, bt = torch.ones(C, device="cuda", dtype=torch.bfloat16), torch.zeros(C, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.groupnorm(x1, gm, bt, silu=True, x2=x2)
    for _ in range(3): fn()
    ts = []
    for _ in range(10):  # single launches, cold L2 (as in the step: the producer's output is mostly evicted or in L2)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    cold = sorted(ts)[len(ts) // 2]
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10): fn()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 10)
    hot = min(ts)
    byt = 2.0 * n * hw * C * 2
    print(f"  {n}x{hw}x{c1}+{c2}: back-to-back {hot * 1e3:6.1f} us ({byt / hot / 1e6 / peak:4.2f} of HBM peak)   single, L2 flushed {cold * 1e3:6.1f} us", flush=True)
