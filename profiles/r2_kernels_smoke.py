"""One small launch of every kernel that was new or re-scheduled in round 2, each checked against fp32 torch math
(written for compute-sanitizer memcheck, which is closed on this pool; useful as a 5-second check of a new build):
stream-K GEMM / conv (all-tiles and hybrid-tail schedules), window-major and cluster GroupNorm, persistent attention
with a split tail and TMA-store epilogue, one-CTA-per-unit attention with a KV-split, conv_in, conv_out route."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvd_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, device="cuda", generator=g) * scale).to(torch.bfloat16)
def check(name, got, ref, tol):
    err = (got.float() - ref.float()).abs().max().item()
    print(f"{name}: max|err| {err:.3e}", flush=True)
    assert err <= tol * max(1.0, ref.float().abs().max().item()), name
F = torch.nn.functional
# stream-K conv, every tile shared (40 tiles) and hybrid tail (160 tiles)
for (n, hw, cin, cout) in [(1, 16, 640, 1280), (2, 32, 320, 640)]:
    x, w = rnd(n, hw, hw, cin), rnd(cout, 9 * cin, scale=(9 * cin) ** -0.5)
    out = ops.conv3x3(x, w)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    check(f"conv stream-K {n}x{hw}x{hw} {cin}->{cout}", out, ref, 3e-2)
a, w = rnd(200, 2560), rnd(640, 2560, scale=2560 ** -0.5)
check("linear stream-K 200x2560x640", ops.linear(a, w), a.float() @ w.float().t(), 3e-2)
# GroupNorm: window-major cluster kernel (8 images) and group-major cluster kernel (1 image)
for (n, hw, c) in [(8, 4096, 320), (1, 1024, 640)]:
    x = rnd(n, hw, c)
    gm, bt = rnd(c, scale=0.2) + 1, rnd(c, scale=0.2)
    ref = F.silu(F.group_norm(x.float().permute(0, 2, 1), 32, gm.float(), bt.float(), 1e-5).permute(0, 2, 1))
    check(f"groupnorm {n}x{hw}x{c}", ops.groupnorm(x, gm, bt, groups=32, eps=1e-5, silu=True), ref, 3e-2)
# attention: persistent (160 units: one wave + a split tail) and one CTA per unit with a KV-split (80 units)
for B in (2, 1):
    H, S = 5, 4096
    q, kv = rnd(B, S, 320), rnd(B, S, 640)
    k, v = kv[:, :, :320], kv[:, :, 320:]
    qh, kh, vh = (t.float().view(B, -1, H, 64).transpose(1, 2) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, S, 320)
    check(f"attention B={B}", ops.attention(q, k, v, H), ref, 2e-2)
# ragged sequence (S_q, S_kv not multiples of 128) through the persistent kernel
B, H, Sq, Skv = 8, 5, 1000, 777
q, kv = rnd(B, Sq, 320), rnd(B, Skv, 640)
qh, kh, vh = (t.float().view(B, -1, H, 64).transpose(1, 2) for t in (q, kv[:, :, :320], kv[:, :, 320:]))
ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, Sq, 320)
check("attention ragged 1000x777", ops.attention(q, kv[:, :, :320], kv[:, :, 320:], H), ref, 2e-2)
# conv_out route: 32-column conv + 4-channel tail
x = rnd(2, 16, 16, 320)
w4 = rnd(4, 320, 3, 3, scale=(9 * 320) ** -0.5)
w32 = torch.zeros(32, 9 * 320, device="cuda", dtype=torch.bfloat16)
w32[:4] = w4.permute(0, 2, 3, 1).reshape(4, -1)
out = ops.head4_to_nchw(ops.conv3x3(x, w32))
check("conv_out via tcgen05 conv", out, F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), padding=1), 2e-2)
torch.cuda.synchronize()
print("r2_kernels_smoke: ok")
