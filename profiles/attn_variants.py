"""Round-2 attention tuning harness: times the attention kernel variants (MVD_ATTN_VARIANT) at the configs[1] top site
(B=8, h=5, S=4096, d=64) and at a view-sharded shape (B=2), checks each against fp32 SDPA, and dumps a clock64 trace of the
softmax warps for the traced variants.  usage: python profiles/attn_variants.py [variant ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mvd_b200 import ops

variants = [int(a) for a in sys.argv[1:]] or [0, 1, 3, 11, 12, 13, 14, 15]
H, S = 5, 4096
C = H * 64
g = torch.Generator(device="cuda").manual_seed(0)


def run(B, variant, reps=20):
    os.environ["MVD_ATTN_VARIANT"] = str(variant)
    qkv = torch.randn(B, S, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(B, S, C, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    for _ in range(3):
        ops.attention(q, k, v, H, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.attention(q, k, v, H, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    def heads(t):
        return t.float().reshape(B, S, H, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(heads(q), heads(k), heads(v)).transpose(1, 2).reshape(B, S, C)
    err = (out.float() - ref).abs().max().item()
    return ms, err


def code(seq, poly, sc, trace=0):
    return 1 + poly + 10 * seq + 100 * sc + 1000 * trace


if not sys.argv[1:]:
    variants = [0, code(0, 0, 1), 5000, 5001, 5002, 5003]
for var in variants:
    for B in (8, 2):
        ms, err = run(B, var)
        print(f"variant {var:4d} B={B}: {ms * 1e3:7.1f} us  {4.0 * S * S * C * B / ms / 1e9:6.0f} TFLOP/s  max|err| {err:.2e}", flush=True)

# traces: clock64 stamps of warp q=0 of both softmax warpgroups, CTA (0,0,0)
names = ["loop top", "S ready", "S in regs", "max (+exchange)", "pv done", "exp done", "P published"]
traced = [5100, 5101]
for var in traced:
    buf = torch.zeros(4 * 8 * 64, dtype=torch.int64, device="cuda")
    os.environ["MVD_ATTN_TRACE_PTR"] = str(buf.data_ptr())
    run(8, var, reps=1)
    os.environ.pop("MVD_ATTN_TRACE_PTR")
    nw = 4 if var >= 4000 else 2
    t = buf.cpu().reshape(4, 64, 8)
    t0 = t[0, 0, 0].item()
    print(f"--- trace variant {var} (cycles; per KV block: deltas between stamps; wg0 = tile A, wg1 = tile B)")
    for wg in range(nw):
        rows = t[wg, 8:24, :7] - t0
        d = rows[:, 1:] - rows[:, :-1]
        period = (rows[1:, 0] - rows[:-1, 0]).float().mean().item()
        nm = ["loop top", "S in regs", "h0 published", "look-ahead S ready", "exps done", "h1 published", "-"] if var >= 5000 else names
        print(f"  wg{wg}: period {period:7.0f} | " + " | ".join(f"{n}: {x:6.0f}" for n, x in zip(nm[1:], d.float().mean(0).tolist())))
        print(f"        exp sections: {[(a, b) for a, b in zip(rows[:5, 4].tolist(), rows[:5, 5].tolist())]}")
