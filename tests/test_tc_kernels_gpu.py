"""GPU parity of the three tensor-core kernels (GEMM/conv/attention) against fp32 torch math on the same
bf16 inputs. Tolerances: outputs are bf16 (rel 2^-8); fp32 accumulation. Written in the test below."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cmp(got, ref, name, atol, rtol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    bad = (err > tol).sum().item()
    print(f"{name}: max_abs={err.max().item():.4e} ref_max={ref.abs().max().item():.3f} cos={cos:.6f} bad={bad}")
    assert bad == 0 and cos > 0.9995, f"{name}: {bad} elements out of tolerance, cos={cos}"


def _randn(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to("cuda", torch.bfloat16)


@pytest.mark.parametrize("tile_n", [0, 64, 128, 160, 256])
@pytest.mark.parametrize("M,N,K", [(1000, 320, 320), (256, 1280, 640), (77, 640, 1024), (4096, 320, 1280),
                                   (640, 640, 320), (512, 1280, 5120)])
def test_linear_plain(M, N, K, tile_n):
    from mvd_b200 import ops

    a = _randn(M, K, seed=1)
    w = _randn(N, K, scale=K ** -0.5, seed=2)
    out = ops.linear(a, w, tile_n=tile_n)
    torch.cuda.synchronize()
    _cmp(out, a.float() @ w.float().t(), f"linear {M}x{N}x{K} bn={tile_n}", atol=2e-2)


def test_linear_epilogues():
    from mvd_b200 import ops

    M, N, K = 2048 + 40, 640, 640
    a = _randn(M, K, seed=1)
    w = _randn(N, K, scale=K ** -0.5, seed=2)
    b = _randn(N, seed=3)
    r = _randn(M, N, seed=4)
    ref = a.float() @ w.float().t() + b.float() + r.float()
    _cmp(ops.linear(a, w, bias=b, residual=r), ref, "linear bias+res", atol=3e-2)
    # strided views: a is a column slice of a wider matrix, out is a column slice too
    wide = _randn(M, 3 * K, seed=5)
    a_v = wide[:, K:2 * K]
    out_wide = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    ops.linear(a_v, w, bias=b, out=out_wide[:, N:])
    _cmp(out_wide[:, N:], a_v.float() @ w.float().t() + b.float(), "linear strided", atol=3e-2)
    assert out_wide[:, :N].abs().max().item() == 0
    # two sources (channel concat)
    a2 = _randn(M, 320, seed=6)
    w2 = _randn(N, K + 320, scale=(K + 320) ** -0.5, seed=7)
    ref = torch.cat([a.float(), a2.float()], 1) @ w2.float().t()
    _cmp(ops.linear(a, w2, a2=a2), ref, "linear two-source", atol=3e-2)
    # per-row-group bias (rows_per_group = 512)
    g = torch.randn(5, N, device="cuda")
    ref = a.float() @ w.float().t() + g[(torch.arange(M, device="cuda") // 512)]
    _cmp(ops.linear(a, w, row_group_bias=g, rows_per_group=512), ref, "linear group-bias", atol=3e-2)


@pytest.mark.parametrize("tile_n", [128, 256])
def test_linear_geglu(tile_n):
    from mvd_b200 import ops

    M, C = 1024 + 8, 320
    a = _randn(M, C, seed=1)
    w = _randn(8 * C, C, scale=C ** -0.5, seed=2)   # rows [0,4C) = value, [4C,8C) = gate
    b = _randn(8 * C, seed=3)
    h = a.float() @ w.float().t() + b.float()
    ref = h[:, :4 * C] * F.gelu(h[:, 4 * C:])
    half = tile_n // 2
    # interleave [value block | gate block] per tile
    wv, wg = w[:4 * C].view(-1, half, C), w[4 * C:].view(-1, half, C)
    wp = torch.cat([wv, wg], 1).reshape(8 * C, C).contiguous()
    bp = torch.cat([b[:4 * C].view(-1, half), b[4 * C:].view(-1, half)], 1).reshape(-1).contiguous()
    out = ops.linear(a, wp, bias=bp, geglu=True, tile_n=tile_n)
    _cmp(out, ref, f"geglu bn={tile_n}", atol=3e-2)


@pytest.mark.parametrize("M,C,N", [(4096, 320, 1280), (1000, 640, 640), (512, 1280, 320)])
def test_linear_row_stats_and_layernorm_fold(M, C, N):
    """Producer epilogue: per-row partial (sum, sum of squares) of the bf16 output, one pair per column tile.
    Consumer: LayerNorm folded into the GEMM (gamma in the weights, mean / rstd applied in the epilogue) against
    F.layer_norm followed by the matmul in fp32 — for a plain linear and for the GEGLU projection."""
    from mvd_b200 import ops
    from mvd_b200.unet import fold_layernorm

    a = _randn(M, C, seed=1)
    w0 = _randn(C, C, scale=C ** -0.5, seed=2)
    b0 = _randn(C, seed=3)
    r = _randn(M, C, scale=2.0, seed=4) + 1.5          # a residual stream with a clear mean
    x, st = ops.linear(a, w0, bias=b0, residual=r, want_stats=True)
    torch.cuda.synchronize()
    xs = x.float()
    got = st.buf.sum(1)
    assert st.buf.shape == (M, st.parts, 2)
    # the statistics are taken on the fp32 values before their bf16 rounding (zero-mean noise, 2^-9 relative per element)
    assert torch.allclose(got[:, 0], xs.sum(1), rtol=0, atol=0.02 * C ** 0.5)
    assert torch.allclose(got[:, 1], (xs * xs).sum(1), rtol=3e-3, atol=0)

    norm = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        norm.weight.copy_(1.0 + 0.3 * torch.randn(C, device="cuda"))
        norm.bias.copy_(0.2 * torch.randn(C, device="cuda"))
        norm.weight.copy_(norm.weight.to(torch.bfloat16).float())
        norm.bias.copy_(norm.bias.to(torch.bfloat16).float())
    w = _randn(N, C, scale=C ** -0.5, seed=5)
    ref = F.layer_norm(xs, (C,), norm.weight, norm.bias, norm.eps) @ w.float().t()
    wg, colsum, cst = fold_layernorm(w, norm)
    out = ops.linear(x, wg, row_group_bias=cst, rows_per_group=M, ln=ops.LNFold(st, colsum, norm.eps))
    _cmp(out, ref, f"layernorm fold {M}x{C}->{N}", atol=3e-2)
    # the un-fused path on the same inputs, for scale: LayerNorm kernel + GEMM
    unf = ops.linear(ops.layernorm(x, norm.weight.to(torch.bfloat16), norm.bias.to(torch.bfloat16), norm.eps), w)
    e_f, e_u = (out.float() - ref).abs().max().item(), (unf.float() - ref).abs().max().item()
    print(f"  max|err| folded {e_f:.3e}  un-fused {e_u:.3e}")
    assert e_f <= 2.0 * e_u + 1e-2

    # GEGLU projection with the fold (interleaved [value | gate] rows per 256-column tile)
    inner, half = 4 * C, 128
    wf = _randn(2 * inner, C, scale=C ** -0.5, seed=6)
    bf = _randn(2 * inner, seed=7)
    hh = F.layer_norm(xs, (C,), norm.weight, norm.bias, norm.eps) @ wf.float().t() + bf.float()
    ref_g = hh[:, :inner] * F.gelu(hh[:, inner:])
    wp = torch.cat([wf[:inner].view(-1, half, C), wf[inner:].view(-1, half, C)], 1).reshape(2 * inner, C).contiguous()
    bp = torch.cat([bf[:inner].view(-1, half), bf[inner:].view(-1, half)], 1).reshape(-1).contiguous()
    wg, colsum, cst = fold_layernorm(wp, norm, bp)
    out_g = ops.linear(x, wg, bias=cst.view(-1).to(torch.bfloat16), geglu=True, tile_n=256,
                       ln=ops.LNFold(st, colsum, norm.eps))
    _cmp(out_g, ref_g, f"layernorm fold + geglu {M}x{C}", atol=8e-2)


@pytest.mark.parametrize("n,hw,cin,cin2,cout", [(1, 8, 1280, 0, 1280), (1, 8, 1280, 1280, 1280), (2, 8, 1280, 1280, 1280),
                                                (1, 16, 1280, 640, 1280), (1, 16, 640, 0, 1280),
                                                (1, 32, 640, 0, 640),      # 80 tiles: every CTA's share spans 2 tiles
                                                (1, 32, 1280, 640, 640),   # the same with two sources
                                                (1, 64, 320, 0, 320),      # 160 tiles: one whole wave + a 12-tile tail
                                                (2, 32, 320, 0, 640),      # 160 tiles, short k-loops
                                                (8, 64, 320, 0, 320),      # 512 wide tiles: 3 whole waves + 68 shared
                                                (4, 32, 640, 640, 640),
                                                (2, 8, 640, 0, 1280)])     # 9 segments per tile at 12-13 k-blocks per CTA
def test_conv3x3_split_k(n, hw, cin, cin2, cout):
    """Launches (or last waves) that do not fill the machine share their k-blocks evenly over the CTAs (stream-K): the
    fp32 partial tiles are summed in k order by the last-arriving CTA. Same result as the fp32 reference, bit-identical
    across launches, ticket counters re-armed; the planner reports the un-split plan without a workspace."""
    from mvd_b200 import ops

    x = _randn(n, hw, hw, cin, seed=1)
    x2 = _randn(n, hw, hw, cin2, seed=5) if cin2 else None
    w9 = _randn(cout, 9 * (cin + cin2), scale=(9 * (cin + cin2)) ** -0.5, seed=2)
    b = _randn(cout, seed=3)
    ib = torch.randn(n, cout, device="cuda")
    r = _randn(n, hw, hw, cout, seed=4)
    outs = [ops.conv3x3(x, w9, bias=b, img_bias=ib, residual=r, x2=x2).clone() for _ in range(3)]
    torch.cuda.synchronize()
    _cmp(outs[0], _conv_ref(x, w9, b, ib, r, 1, x2=x2), f"conv split-K {n}x{hw}x{hw} {cin}+{cin2}->{cout}", atol=4e-2)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("M,K,N", [(64, 5120, 1280),     # ff.net[2] at the 8x8 level of one sample
                                   (4096, 1280, 320),    # ff.net[2] at 64x64 of one sample: 160 tiles, 20 k-blocks
                                   (1024, 2560, 640),    # 80 tiles, 40 k-blocks
                                   (32768, 1280, 320),   # 512 wide tiles: whole waves + shared tail
                                   (200, 1280, 1280)])   # ragged rows
def test_linear_split_k_long_k(M, K, N):
    from mvd_b200 import ops

    a, w, b, r = _randn(M, K, seed=1), _randn(N, K, scale=K ** -0.5, seed=2), _randn(N, seed=3), _randn(M, N, seed=4)
    outs = [ops.linear(a, w, bias=b, residual=r).clone() for _ in range(2)]
    _cmp(outs[0], a.float() @ w.float().t() + b.float() + r.float(), f"linear stream-K {M}x{K}x{N}", atol=3e-2)
    assert torch.equal(outs[0], outs[1])


def test_stream_k_shape_sweep():
    """Every CTA count / segment layout the planner can produce for small launches: M, N, K swept so that the number
    of tiles, k-blocks per CTA and segments per tile all vary (a wrong slot bound corrupts the neighbouring tile)."""
    from mvd_b200 import ops

    bad = []
    for M in (64, 128, 200, 384, 640, 1024, 1536):
        for N in (320, 640, 1280):
            for K in (640, 1280, 2560, 5760):
                a, w = _randn(M, K, seed=M + K), _randn(N, K, scale=K ** -0.5, seed=N + K)
                out = ops.linear(a, w)
                ref = a.float() @ w.float().t()
                err = (out.float() - ref).abs().max().item()
                if not err <= 3e-2 * max(1.0, ref.abs().max().item()):
                    bad.append((M, N, K, err))
    assert not bad, bad


def test_stream_k_matches_whole_tile_schedule(monkeypatch):
    """The same launch with and without the workspace (stream-K on / off): both within bf16 rounding of each other."""
    from mvd_b200 import ops

    x = _randn(1, 32, 32, 640, seed=1)
    w9 = _randn(640, 9 * 640, scale=(9 * 640) ** -0.5, seed=2)
    b = _randn(640, seed=3)
    on = ops.conv3x3(x, w9, bias=b)
    monkeypatch.setattr(ops, "SPLIT_K", False)
    off = ops.conv3x3(x, w9, bias=b)
    d = (on.float() - off.float()).abs().max().item()
    assert d <= 2e-2 * max(1.0, off.float().abs().max().item()), d  # one bf16 rounding step of the largest value


def test_film_epilogue_linear_and_conv():
    """out * scale[g] + shift[g] fused behind bias / residual: per row group (linear) and per image (conv)."""
    from mvd_b200 import ops

    n, hw, c = 4, 16, 320
    a = _randn(n * hw * hw, c, seed=1)
    w = _randn(c, c, scale=c ** -0.5, seed=2)
    b = _randn(c, seed=3)
    r = _randn(n * hw * hw, c, seed=4)
    scale = 1.0 + 0.5 * torch.randn(n, c, device="cuda")
    shift = 0.3 * torch.randn(n, c, device="cuda")
    ref = (a.float() @ w.float().t() + b.float() + r.float()).view(n, hw * hw, c) * scale[:, None] + shift[:, None]
    out = ops.linear(a, w, bias=b, residual=r, rows_per_group=hw * hw, film=(scale, shift))
    _cmp(out, ref.view(-1, c), "linear + film", atol=4e-2)
    x = _randn(n, hw, hw, c, seed=5)
    w9 = _randn(c, 9 * c, scale=(9 * c) ** -0.5, seed=6)
    refc = _conv_ref(x, w9, b, None, r.view(n, hw, hw, c), 1) * scale[:, None, None] + shift[:, None, None]
    outc = ops.conv3x3(x, w9, bias=b, residual=r.view(n, hw, hw, c), film=(scale, shift))
    _cmp(outc, refc, "conv + film", atol=4e-2)
    outs = ops.conv3x3(x, w9, bias=b, stride=2, film=(scale, shift))
    _cmp(outs, _conv_ref(x, w9, b, None, None, 2) * scale[:, None, None] + shift[:, None, None], "conv s2 + film", atol=4e-2)


def _conv_ref(x, w9, bias, img_bias, residual, stride, x2=None):
    xin = x.float() if x2 is None else torch.cat([x.float(), x2.float()], -1)
    cin = xin.shape[-1]
    cout = w9.shape[0]
    w = w9.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    y = F.conv2d(xin.permute(0, 3, 1, 2), w, bias=None if bias is None else bias.float(), stride=stride, padding=1)
    y = y.permute(0, 2, 3, 1)
    if img_bias is not None:
        y = y + img_bias[:, None, None, :]
    if residual is not None:
        y = y + residual.float()
    return y


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 64), (2, 64, 64, 320, 320), (3, 8, 8, 1280, 1280),
                                             (1, 32, 32, 640, 320), (2, 24, 24, 128, 64), (1, 96, 96, 64, 64),
                                             (5, 12, 12, 64, 96)])
def test_conv3x3(n, h, w, cin, cout):
    from mvd_b200 import ops

    x = _randn(n, h, w, cin, seed=1)
    w9 = _randn(cout, 9 * cin, scale=(9 * cin) ** -0.5, seed=2)
    b = _randn(cout, seed=3)
    ib = torch.randn(n, cout, device="cuda")
    r = _randn(n, h, w, cout, seed=4)
    _cmp(ops.conv3x3(x, w9), _conv_ref(x, w9, None, None, None, 1), f"conv {n}x{h}x{w} {cin}->{cout}", atol=3e-2)
    _cmp(ops.conv3x3(x, w9, bias=b, img_bias=ib, residual=r), _conv_ref(x, w9, b, ib, r, 1),
         f"conv+epi {n}x{h}x{w} {cin}->{cout}", atol=4e-2)


@pytest.mark.parametrize("n,h,w,c", [(2, 64, 64, 320), (3, 16, 16, 1280), (1, 32, 32, 640), (2, 24, 24, 64)])
def test_conv3x3_stride2(n, h, w, c):
    from mvd_b200 import ops

    x = _randn(n, h, w, c, seed=1)
    w9 = _randn(c, 9 * c, scale=(9 * c) ** -0.5, seed=2)
    b = _randn(c, seed=3)
    _cmp(ops.conv3x3(x, w9, bias=b, stride=2), _conv_ref(x, w9, b, None, None, 2), f"conv s2 {n}x{h}x{w} {c}",
         atol=3e-2)


@pytest.mark.parametrize("n,hw,c1,c2,cout", [(2, 32, 640, 320, 640), (2, 16, 1280, 1280, 1280), (2, 8, 1280, 1280, 1280),
                                              (1, 16, 1280, 640, 1280)])
def test_conv3x3_two_source(n, hw, c1, c2, cout):
    """Channel-concatenated skip connections of the up path (two tensor maps, one k-loop), incl. the 1280+1280 sites
    of up_blocks[0..1] at the 16x16 / 8x8 levels and the 1280+640 one."""
    from mvd_b200 import ops

    x = _randn(n, hw, hw, c1, seed=1)
    x2 = _randn(n, hw, hw, c2, seed=5)
    w9 = _randn(cout, 9 * (c1 + c2), scale=(9 * (c1 + c2)) ** -0.5, seed=2)
    b = _randn(cout, seed=3)
    _cmp(ops.conv3x3(x, w9, bias=b, x2=x2), _conv_ref(x, w9, b, None, None, 1, x2=x2),
         f"conv two-source {c1}+{c2}->{cout} @{hw}", atol=3e-2)


def _attn_ref(q, k, v, heads, scale):
    B, Sq, C = q.shape
    qh = q.float().view(B, Sq, heads, 64).transpose(1, 2)
    kh = k.float().view(B, -1, heads, 64).transpose(1, 2)
    vh = v.float().view(B, -1, heads, 64).transpose(1, 2)
    o = F.scaled_dot_product_attention(qh, kh, vh, scale=scale)
    return o.transpose(1, 2).reshape(B, Sq, C)


@pytest.mark.parametrize("B,heads,Sq,Skv", [(2, 5, 256, 256), (2, 20, 64, 77), (1, 10, 200, 1000), (4, 5, 4096, 4096),
                                             (1, 5, 1024, 4 * 1024), (8, 20, 64, 64), (8, 5, 2000, 1000), (8, 10, 1024, 1024),
                                             (8, 5, 4096, 4 * 4096 + 8)])
@pytest.mark.parametrize("qscale", [1.0, 6.0])
def test_attention(B, heads, Sq, Skv, qscale):
    from mvd_b200 import ops

    C = heads * 64
    q = _randn(B, Sq, C, scale=qscale, seed=1)
    k = _randn(B, Skv, C, seed=2)
    v = _randn(B, Skv, C, seed=3)
    out = ops.attention(q, k, v, heads)
    torch.cuda.synchronize()
    _cmp(out, _attn_ref(q, k, v, heads, 0.125), f"attn B{B} h{heads} {Sq}x{Skv} qs{qscale}", atol=1.5e-2)


@pytest.mark.parametrize("B,heads,Sq,Skv", [(2, 5, 4096, 4096),     # 160 units on 148 SMs: 12 tail units x 8 KV parts
                                             (8, 5, 4096, 4096),     # 640 units: 48 tail units x 3 parts
                                             (3, 5, 4000, 1000 + 8),  # ragged rows and a masked last KV block in a part
                                             (1, 10, 4096, 2048),    # 160 units, 16 KV blocks: 4 parts
                                             (2, 10, 2304, 2304)])   # 96^2 latent level-1 site of a view-sharded rank
def test_attention_kv_split_tail(B, heads, Sq, Skv, monkeypatch):
    """The last partial wave is split along S_kv and merged by the last-arriving CTA: same result as the un-split
    launch (MVD_ATTN_SPLIT=0 is read once per process, so the comparison is against fp32 SDPA), bit-identical
    across repeated launches (fixed merge order), ticket counters re-armed."""
    from mvd_b200 import ops

    C = heads * 64
    q = _randn(B, Sq, C, scale=2.0, seed=1)
    k = _randn(B, Skv, C, seed=2)
    v = _randn(B, Skv, C, seed=3)
    outs = [ops.attention(q, k, v, heads).clone() for _ in range(3)]
    torch.cuda.synchronize()
    _cmp(outs[0], _attn_ref(q, k, v, heads, 0.125), f"attn split B{B} h{heads} {Sq}x{Skv}", atol=1.5e-2)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_attention_shared_kv_batch_broadcast():
    """k/v with batch stride 0 (one reference sequence for every view, configs[3] cross-view mode) == the same K/V
    materialised per batch entry, bit for bit; both kernels (one-tile and two-tile)."""
    from mvd_b200 import ops

    for (B, heads, Sq, Skv) in [(4, 5, 256, 1024), (8, 5, 2304, 4 * 2304)]:
        C = heads * 64
        q = _randn(B, Sq, C, seed=1)
        k1, v1 = _randn(1, Skv, C, seed=2), _randn(1, Skv, C, seed=3)
        shared = ops.attention(q, k1.expand(B, -1, -1), v1.expand(B, -1, -1), heads)
        full = ops.attention(q, k1.repeat(B, 1, 1), v1.repeat(B, 1, 1), heads)
        torch.cuda.synchronize()
        assert torch.equal(shared, full)
        _cmp(shared, _attn_ref(q, k1.repeat(B, 1, 1), v1.repeat(B, 1, 1), heads, 0.125), "attn shared kv", atol=1.5e-2)


def test_attention_strided_fused_qkv():
    """q/k/v as column slices of one fused projection output, out written into a slice of a wider buffer."""
    from mvd_b200 import ops

    B, S, heads = 2, 384, 5
    C = heads * 64
    qkv = _randn(B, S, 3 * C, seed=1)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    cat = torch.zeros(B, S, 2 * C, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, heads, out=cat[:, :, C:])
    _cmp(cat[:, :, C:], _attn_ref(q.contiguous(), k.contiguous(), v.contiguous(), heads, 0.125), "attn strided",
         atol=1.5e-2)
    assert cat[:, :, :C].abs().max().item() == 0


def test_attention_config3_cross_view_length():
    """configs[3] top site: one 96x96 view (9216 queries) attending over all 8 views' tokens (S_kv = 73,728), 5 heads.
    Reference computed in fp32 by query chunks (the full score matrix would be 13.6 GB)."""
    from mvd_b200 import ops

    heads, Sq, Skv = 5, 9216, 8 * 9216
    C = heads * 64
    q, k, v = _randn(1, Sq, C, seed=1), _randn(1, Skv, C, seed=2), _randn(1, Skv, C, seed=3)
    out = ops.attention(q, k, v, heads)
    torch.cuda.synchronize()
    kh = k.float().view(1, Skv, heads, 64).transpose(1, 2)
    vh = v.float().view(1, Skv, heads, 64).transpose(1, 2)
    ref = torch.empty(1, Sq, C, device="cuda")
    for s in range(0, Sq, 1024):
        qh = q[:, s:s + 1024].float().view(1, -1, heads, 64).transpose(1, 2)
        p = torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, dim=-1)
        ref[:, s:s + 1024] = (p @ vh).transpose(1, 2).reshape(1, -1, C)
    _cmp(out, ref, "attn 9216 x 73728 (configs[3])", atol=2e-3)
