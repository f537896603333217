import os, sys
sys.path.insert(0, "/root/repo")
import torch
from mvd_b200 import ops
B, H, S = 8, 5, 4096
C = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, S, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(B, S, C, device="cuda", dtype=torch.bfloat16)
os.environ["MVD_ATTN_TRACE"] = "1"
for _ in range(3):
    ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, out=out)
torch.cuda.synchronize()
os.environ["MVD_ATTN_TRACE"] = "2"
ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, out=out)
torch.cuda.synchronize()
