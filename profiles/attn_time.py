import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
B, H, S = 8, 5, 4096
C = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, S, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(B, S, C, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, out=out)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, out=out)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = min(ts)
print(f"debug={os.environ.get('MVD_ATTN_DEBUG','0')}: {ms*1e3:.1f} us  {4.0*S*S*C*B/ms/1e9:.0f} TFLOP/s")
