"""Round-2 attention harness: times the two-tile attention kernel at the configs[1] top site (B=8, h=5, S=4096, d=64), at
the view-sharded shapes (B=2, B=1) and at the 32x32 level, with the KV-split tail on and off and for the polynomial
exp2 shares, checks each against fp32 SDPA, and dumps a clock64 trace of the softmax warps.
usage: python profiles/attn_variants.py          (each configuration runs in a fresh process: the switches are read once)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.nn.functional as F
from mvd_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
def run(B, H, S, reps=20):
    C = H * 64
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
    hd = lambda t: t.float().reshape(B, S, H, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(hd(q), hd(k), hd(v)).transpose(1, 2).reshape(B, S, C)
    err = (out.float() - ref).abs().max().item()
    print(f"  B={B} h={H} S={S}: {ms * 1e3:7.1f} us  {4.0 * S * S * C * B / ms / 1e9:6.0f} TFLOP/s  max|err| {err:.2e}", flush=True)
if os.environ.get("TRACE") == "1":
    buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
    os.environ["MVD_ATTN_TRACE_PTR"] = str(buf.data_ptr())
    for B in (2, 8):
        buf.zero_()
        run(B, 5, 4096, reps=1)
        c = buf.cpu()[1024:1032].tolist()
        print(f"  last CTA of the grid at B={B}: {c[6]} KV blocks, {c[7]} parts; cycles: main loop {c[1] - c[0]}, wait O {c[2] - c[1]}, "
              f"partial write + fence {c[3] - c[2]}, ticket + merge {c[4] - c[3]}")
    t = buf.cpu()[:1024].reshape(2, 64, 8)
    names = ["S ready", "S in regs (+PV(j-1) done)", "row max", "O rescale", "exp + P stored", "P published"]
    for wg in (0, 1):
        rows = t[wg, 8:24, :7] - t[0, 0, 0]
        d = (rows[:, 1:] - rows[:, :-1]).float().mean(0).tolist()
        period = (rows[1:, 0] - rows[:-1, 0]).float().mean().item()
        print(f"  trace wg{wg}: period {period:6.0f} | " + " | ".join(f"{n}: {x:5.0f}" for n, x in zip(names, d)))
else:
    for (B, H, S) in [(8, 5, 4096), (2, 5, 4096), (1, 5, 4096), (8, 10, 1024), (2, 10, 1024)]:
        run(B, H, S)
""" % ROOT

for label, env in [("default (KV-split tail on, MUFU only)", {}), ("MVD_ATTN_SPLIT=0", {"MVD_ATTN_SPLIT": "0"}),
                   ("MVD_ATTN_POLY8=1", {"MVD_ATTN_POLY8": "1"}), ("MVD_ATTN_POLY8=2", {"MVD_ATTN_POLY8": "2"}),
                   ("trace", {"TRACE": "1"})]:
    print(f"--- {label}", flush=True)
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, "-c", CHILD], env=e, check=False)
