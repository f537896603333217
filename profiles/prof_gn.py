import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(8, 4096, 320, device="cuda", generator=g).to(torch.bfloat16)
gm = torch.ones(320, device="cuda", dtype=torch.bfloat16); bt = torch.zeros(320, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.groupnorm(x, gm, bt, silu=True)
torch.cuda.synchronize()
