"""Attention kernel time, back to back (CUDA graph of 10 launches, min of 5 replays) and as single flushed launches,
for the shapes of the step. Run once per setting of MVD_ATTN_PERSIST / MVD_ATTN_SPLIT (read once per process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
print(f"MVD_ATTN_PERSIST={os.environ.get('MVD_ATTN_PERSIST', '1')} MVD_ATTN_SPLIT={os.environ.get('MVD_ATTN_SPLIT', '1')}")
for (B, H, S, Skv) in [(8, 5, 4096, 4096), (8, 10, 1024, 1024), (4, 5, 4096, 4096), (2, 5, 4096, 4096), (1, 5, 4096, 4096),
                       (1, 10, 1024, 1024), (2, 10, 1024, 1024), (2, 5, 9216, 73728)]:
    C = H * 64
    q = torch.randn(B, S, C, device="cuda", generator=g).to(torch.bfloat16)
    kv = torch.randn(B, Skv, 2 * C, device="cuda", generator=g).to(torch.bfloat16)
    k, v = kv[:, :, :C], kv[:, :, C:]
    out = torch.empty(B, S, C, device="cuda", dtype=torch.bfloat16)
    ref = None
    if Skv <= 4096:
        qh, kh, vh = (t.float().view(B, -1, H, 64).transpose(1, 2) for t in (q, k, v))
        ref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, S, C)
    fn = lambda: ops.attention(q, k, v, H, out=out)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item() if ref is not None else float("nan")
    reps = 10 if Skv <= 4096 else 2
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps): fn()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps)
    ms = min(ts)
    print(f"  B={B} h={H} S={S} Skv={Skv}: {ms * 1e3:8.1f} us  {4.0 * S * Skv * C * B / ms / 1e9:6.0f} TFLOP/s  max|err| {err:.2e}", flush=True)
