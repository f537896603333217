"""clock64 stamps of the softmax warps of CTA 0 (non-persistent attn_pair_kernel, trace build selected by
MVD_ATTN_TRACE_PTR): per KV block and tile, where the cycles go, and the phase between the two warpgroups.
stamps per block: 0 loop top | 1 S ready | 2 S in registers, PV(j-1) done, S handed back | 3 row max | 4 rescale |
5 exponentials + P stored | 6 P published"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
B, H, S = 8, 5, 4096
C = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, S, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(B, S, C, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
for _ in range(3):
    ops.attention(q, k, v, H, out=out)
torch.cuda.synchronize()
trace = torch.zeros(2048, device="cuda", dtype=torch.int64)
os.environ["MVD_ATTN_TRACE_PTR"] = str(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention(q, k, v, H, out=out, split_tail=False)
e1.record()
torch.cuda.synchronize()
print(f"traced launch: {e0.elapsed_time(e1) * 1e3:.1f} us")
t = trace.cpu().view(-1).tolist()
if os.environ.get("MVD_ATTN_PERSIST", "1") != "0":
    # persistent kernel: item-boundary stamps of CTA 0 (0 item start, 1 first S ready, 2 block loop done, 3 O final, 4 stored)
    base = t[0]
    print("tile item   start  S-wait   loop   O-wait  store | item total   gap to next item's first S")
    for it in range(8):
        for tl in range(2):
            s = [t[(tl * 8 + it) * 8 + i] for i in range(5)]
            if s[0] == 0:
                continue
            nxt = t[(tl * 8 + it + 1) * 8 + 1] if it < 7 else 0
            print(f"  {'AB'[tl]}  {it:3d} {s[0]-base:8d} {s[1]-s[0]:6d} {s[2]-s[1]:7d} {s[3]-s[2]:7d} {s[4]-s[3]:6d} | {s[4]-s[0]:8d}   "
                  f"{(nxt - s[2]) if nxt else 0:8d}")
    sys.exit(0)
nb = S // 128
base = t[0]
print("tile blk   top  s_wait  ld+pv  max  resc   exp   pub | period   A->B lag")
prev = [None, None]
for j in range(nb):
    for tl in range(2):
        s = [t[tl * 512 + j * 8 + i] for i in range(7)]
        per = s[0] - prev[tl] if prev[tl] is not None else 0
        prev[tl] = s[0]
        lag = t[512 + j * 8 + 1] - t[j * 8 + 1] if tl == 1 else 0
        print(f"  {'AB'[tl]}  {j:3d} {s[0]-base:6d} {s[1]-s[0]:6d} {s[2]-s[1]:6d} {s[3]-s[2]:5d} {s[4]-s[3]:5d} {s[5]-s[4]:5d} {s[6]-s[5]:5d} | {per:6d} {lag:8d}")
print("CTA-level (last CTA): start", t[1024] - t[1024], "loop end", t[1025] - t[1024], "O final", t[1026] - t[1024], "stored", t[1028] - t[1024],
      "blocks", t[1030], "parts", t[1031])
print("CTA 0 total loop cycles (tile A):", t[(nb - 1) * 8 + 6] - t[0])
