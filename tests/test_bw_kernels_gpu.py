"""GPU parity of the bandwidth-bound kernels against fp32 torch math on the same inputs.
Tolerance: bf16 output rounding (rel 2^-8) on O(1) values -> atol 2e-2 unless stated."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _r(*shape, scale=1.0, seed=0, dtype=torch.bfloat16, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale + shift).to("cuda", dtype)


def _close(got, ref, name, atol=2e-2, rtol=1e-2):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    bad = (err > atol + rtol * ref.abs()).sum().item()
    print(f"{name}: max_abs={err.max().item():.3e} ref_max={ref.abs().max().item():.3f} bad={bad}")
    assert bad == 0, name


@pytest.mark.parametrize("n,hw,c1,c2", [(2, 4096, 320, 0), (3, 64, 1280, 1280), (2, 1024, 640, 320), (1, 256, 1280, 640),
                                        (2, 9216, 320, 0), (1, 37, 64, 0), (1, 4096, 640, 320), (1, 16384, 320, 0), (1, 16384, 1280, 0),
                                        # full batches: the row-major cluster kernel (and its fall-backs when an image's
                                        # rows do not fit the cluster's shared memory)
                                        (8, 4096, 320, 0), (8, 4096, 320, 320), (8, 1024, 640, 0), (8, 1024, 1280, 640),
                                        (8, 1024, 640, 320), (8, 256, 1280, 1280), (8, 256, 1280, 0), (6, 1000, 320, 0),
                                        (16, 1024, 640, 0), (8, 9216, 320, 0)])
@pytest.mark.parametrize("silu", [False, True])
def test_groupnorm(n, hw, c1, c2, silu):
    from mvd_b200 import ops

    x1 = _r(n, hw, c1, seed=1, shift=0.3)
    x2 = _r(n, hw, c2, seed=2, scale=2.0) if c2 else None
    C = c1 + c2
    gamma, beta = _r(C, seed=3, shift=1.0, scale=0.2), _r(C, seed=4, scale=0.2)
    out = ops.groupnorm(x1, gamma, beta, groups=32, eps=1e-5, silu=silu, x2=x2)
    xin = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], -1)
    ref = F.group_norm(xin.permute(0, 2, 1), 32, gamma.float(), beta.float(), 1e-5).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    _close(out, ref, f"gn n{n} hw{hw} c{c1}+{c2} silu{silu}", atol=2.5e-2)


@pytest.mark.parametrize("M,C", [(1000, 320), (513, 640), (64, 1280), (7, 2048)])
def test_layernorm(M, C):
    from mvd_b200 import ops

    x = _r(M, C, seed=1, shift=0.5, scale=1.5)
    g, b = _r(C, seed=2, shift=1.0, scale=0.1), _r(C, seed=3, scale=0.1)
    _close(ops.layernorm(x, g, b), F.layer_norm(x.float(), (C,), g.float(), b.float(), 1e-5), f"ln {M}x{C}")
    # strided rows
    wide = _r(M, 2 * C, seed=4)
    _close(ops.layernorm(wide[:, C:], g, b), F.layer_norm(wide[:, C:].float(), (C,), g.float(), b.float(), 1e-5),
           f"ln strided {M}x{C}")


def test_layernorm_f32():
    from mvd_b200 import ops

    x = _r(4, 512, seed=1, dtype=torch.float32, scale=3.0)
    g, b = _r(512, seed=2, shift=1.0, scale=0.1), _r(512, seed=3, scale=0.1)
    ref = F.silu(F.layer_norm(x, (512,), g.float(), b.float(), 1e-5))
    _close(ops.layernorm_f32(x, g, b, silu=True), ref, "ln f32", atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("B,S,C", [(4, 4096, 320), (4, 64, 1280), (2, 1024, 640), (1, 256, 64)])
def test_refnorm_pixel(B, S, C):
    """4-D reference features [B,C,H,W]: stats over (batch, channel) per pixel, unbiased std (attention.py:95-103)."""
    from mvd_b200 import ops

    x = _r(B, S, C, seed=1, shift=0.2, scale=1.7)
    nchw = x.float().permute(0, 2, 1)  # [B, C, S] stands for [B,C,H,W]
    r = nchw - nchw.mean(dim=(0, 1), keepdim=True)
    r = r / torch.clamp(r.std(dim=(0, 1), keepdim=True), min=1e-6) * 0.5
    _close(ops.refnorm(x, per_pixel=True), r.permute(0, 2, 1), f"refnorm pixel {B}x{S}x{C}", atol=1e-2)


@pytest.mark.parametrize("B,S,C", [(4, 4096, 320), (2, 77, 1280), (1, 16384, 320)])
def test_refnorm_channel(B, S, C):
    """3-D reference features [B,S,C]: stats over (batch, sequence) per channel."""
    from mvd_b200 import ops

    x = _r(B, S, C, seed=1, shift=-0.4, scale=0.7)
    xf = x.float()
    r = xf - xf.mean(dim=(0, 1), keepdim=True)
    r = r / torch.clamp(r.std(dim=(0, 1), keepdim=True), min=1e-6) * 0.5
    _close(ops.refnorm(x, per_pixel=False), r, f"refnorm channel {B}x{S}x{C}", atol=1e-2)


@pytest.mark.parametrize("S,C,rep", [(4 * 256, 320, 8), (77, 64, 2), (2048, 1280, 16)])
def test_refnorm_replicated_reference(S, C, rep):
    """One copy [1,S,C] standing for `rep` identical batch entries (cross-view mode): same result as normalising the
    materialised [rep,S,C] tensor the way attention.py:95-103 does (mean over (0,1), UNBIASED std over rep*S values)."""
    from mvd_b200 import ops

    x = _r(1, S, C, seed=1, shift=0.3, scale=0.9)
    xf = x.float().repeat(rep, 1, 1)
    r = xf - xf.mean(dim=(0, 1), keepdim=True)
    r = r / torch.clamp(r.std(dim=(0, 1), keepdim=True), min=1e-6) * 0.5
    got = ops.refnorm(x, per_pixel=False, replication=rep)
    assert got.shape == (1, S, C)
    _close(got, r[:1], f"refnorm replicated x{rep} {S}x{C}", atol=1e-2)
    with pytest.raises(Exception):
        ops.refnorm(x, per_pixel=True, replication=rep)


def test_refnorm_constant_input_clamps():
    from mvd_b200 import ops

    x = torch.full((2, 16, 64), 1.5, device="cuda", dtype=torch.bfloat16)
    assert ops.refnorm(x, per_pixel=True).abs().max().item() == 0
    assert ops.refnorm(x, per_pixel=False).abs().max().item() == 0


@pytest.mark.parametrize("n,V,hw,C", [(8, 4, 4096, 320), (4, 4, 64, 1280), (2, 1, 1024, 640)])
def test_film(n, V, hw, C):
    from mvd_b200 import ops

    x = _r(n, hw, C, seed=1)
    mod = _r(V, 2 * C, seed=2, dtype=torch.float32)
    s = 0.7
    idx = torch.arange(n, device="cuda") % V
    scale = (torch.sigmoid(mod[:, :C]) * 2 * s)[idx][:, None, :]
    shift = (mod[:, C:] * s)[idx][:, None, :]
    _close(ops.film(x, mod, s), x.float() * scale + shift, f"film {n}x{hw}x{C}")


@pytest.mark.parametrize("M,N,K", [(8, 1280, 320), (4, 512, 9), (4, 1024, 1020), (16, 10240, 1280), (1, 8, 512)])
def test_small_linear(M, N, K):
    from mvd_b200 import ops

    x = _r(M, K, seed=1, dtype=torch.float32)
    w = _r(N, K, seed=2, scale=K ** -0.5)
    b = _r(N, seed=3)
    ref = F.silu(F.silu(x) @ w.float().t() + b.float())
    _close(ops.small_linear(x, w, b, silu_in=True, silu_out=True), ref, f"small_linear {M}x{N}x{K}", atol=1e-3,
           rtol=1e-3)
    _close(ops.small_linear(x, w), x @ w.float().t(), f"small_linear plain {M}x{N}x{K}", atol=1e-3, rtol=1e-3)


def test_timestep_embedding():
    from mvd_b200 import ops

    t = torch.tensor([981.0], device="cuda")
    out = ops.timestep_embedding(t, batch=8)
    half = 160
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
    ang = t[:, None] * freqs[None]
    ref = torch.cat([torch.cos(ang), torch.sin(ang)], -1).expand(8, -1)
    _close(out, ref, "timestep", atol=2e-3, rtol=0)
    t8 = torch.arange(8, device="cuda", dtype=torch.float32) * 100 + 1
    ang = t8[:, None] * freqs[None]
    _close(ops.timestep_embedding(t8, batch=8), torch.cat([torch.cos(ang), torch.sin(ang)], -1), "timestep per-sample",
           atol=2e-3, rtol=0)


def test_camera_front():
    from mvd_b200 import ops

    g = torch.Generator().manual_seed(0)
    src = torch.randn(4, 3, 4, generator=g).cuda()
    tgt = torch.randn(4, 3, 4, generator=g).cuda()
    r, enc, t_rel = ops.camera_front(src, tgt, 170, 10.0)
    R = torch.bmm(tgt[:, :, :3], src[:, :, :3].transpose(1, 2))
    T = tgt[:, :, 3] - torch.bmm(R, src[:, :, 3:]).squeeze(2)
    freqs = torch.exp(torch.linspace(0.0, math.log(10.0), 170, device="cuda"))
    ang = T[:, :, None] * freqs[None, None]
    ref = torch.cat([torch.sin(ang), torch.cos(ang)], -1).reshape(4, -1)
    _close(r, R.reshape(4, 9), "camera R", atol=1e-5, rtol=1e-5)
    _close(t_rel, T, "camera T", atol=1e-5, rtol=1e-5)
    _close(enc, ref, "camera posenc", atol=2e-4, rtol=0)


def test_conv_in_out():
    from mvd_b200 import ops

    lat = _r(4, 4, 64, 64, seed=1, dtype=torch.float32)
    w = _r(320, 4, 3, 3, seed=2, scale=1 / 6)
    b = _r(320, seed=3, scale=0.1)
    mod = _r(4, 8, seed=4, dtype=torch.float32)
    wk = w.permute(0, 2, 3, 1).contiguous()
    out = ops.conv_in(lat, wk, b, n_img=8, mod=mod, strength=1.0)
    scale = torch.sigmoid(mod[:, :4]) * 2
    x = lat * scale[:, :, None, None] + mod[:, 4:, None, None]
    ref = F.conv2d(x.repeat(2, 1, 1, 1), w.float(), b.float(), padding=1).permute(0, 2, 3, 1)
    _close(out, ref, "conv_in film cfg", atol=2e-2)
    out = ops.conv_in(lat, wk, b, n_img=4)
    _close(out, F.conv2d(lat, w.float(), b.float(), padding=1).permute(0, 2, 3, 1), "conv_in plain", atol=2e-2)

    x = _r(3, 32, 32, 320, seed=5)
    w = _r(4, 320, 3, 3, seed=6, scale=(9 * 320) ** -0.5)
    b = _r(4, seed=7, scale=0.1)
    out = ops.conv_out(x, w.permute(0, 2, 3, 1).contiguous(), b)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float(), padding=1)
    _close(out, ref, "conv_out", atol=1e-3, rtol=1e-3)
    # the route the model takes: 32-column tcgen05 conv (4 weight rows zero-padded) + the 4-channel fp32 NCHW tail
    w32 = torch.zeros(32, 9 * 320, device="cuda", dtype=torch.bfloat16)
    w32[:4] = w.permute(0, 2, 3, 1).reshape(4, -1)
    b32 = torch.zeros(32, device="cuda", dtype=torch.bfloat16)
    b32[:4] = b
    full = ops.conv3x3(x, w32, bias=b32)
    assert full[..., 4:].abs().max().item() == 0.0
    out_tc = ops.head4_to_nchw(full)
    assert out_tc.shape == ref.shape and out_tc.dtype == torch.float32
    _close(out_tc, ref, "conv_out via tcgen05 conv", atol=1.6e-2, rtol=8e-3)  # one bf16 rounding of the result


def test_layout_misc():
    from mvd_b200 import ops

    x = _r(2, 16, 16, 64, seed=1)
    up = ops.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    a, b = _r(1000, 64, seed=2), _r(1000, 64, seed=3)
    _close(ops.add(a, b), a.float() + b.float(), "add")
    f = _r(3, 77, 1024, seed=4, dtype=torch.float32)
    assert torch.equal(ops.cast_bf16(f), f.to(torch.bfloat16))
    nchw = _r(2, 320, 24 * 24, seed=5, dtype=torch.float32)
    assert torch.equal(ops.transpose_batched(nchw), nchw.transpose(1, 2).to(torch.bfloat16))
    nhwc = _r(2, 100, 72, seed=6)
    assert torch.equal(ops.transpose_batched(nhwc, out_dtype=torch.float32), nhwc.transpose(1, 2).float())


def test_cfg_ddpm_step():
    from mvd_b200 import ops

    lat = _r(4, 4, 64, 64, seed=1, dtype=torch.float32)
    mo = _r(8, 4, 64, 64, seed=2, dtype=torch.float32)
    noise = _r(4, 4, 64, 64, seed=3, dtype=torch.float32)
    g, sa, sb, c0, ct, sg = 3.0, 0.8, 0.6, 0.3, 0.65, 0.1
    v = mo[:4] + g * (mo[4:] - mo[:4])
    x0 = sa * lat - sb * v
    ref = c0 * x0 + ct * lat + sg * noise
    out = ops.cfg_ddpm_step(mo, lat.clone(), noise, 2, g, sa, sb, c0, ct, sg)
    _close(out, ref, "cfg+ddpm", atol=1e-5, rtol=1e-5)
    out = ops.cfg_ddpm_step(mo[:4].contiguous(), lat.clone(), None, 1, 1.0, sa, sb, c0, ct, 0.0)
    _close(out, c0 * (sa * lat - sb * mo[:4]) + ct * lat, "ddpm no-cfg", atol=1e-5, rtol=1e-5)


def test_cfg_ddpm_step_table_matches_scalar_variant():
    """Device-table step (graph replay) == scalar-argument step, for every step of a 6-step schedule."""
    import mvd_b200
    from mvd_b200 import ops

    sched = mvd_b200.ShiftSNRScheduler.from_scheduler(mvd_b200.DDPMScheduler(), shift_mode="interpolated",
                                                      shift_scale=6.0, scheduler_class=mvd_b200.DDPMScheduler)
    sched.set_timesteps(6)
    ts = [int(t) for t in sched.timesteps]
    coef = torch.zeros(len(ts), 8)
    for i, t in enumerate(ts):
        coef[i, 0] = float(t)
        coef[i, 1:6] = torch.tensor(sched.coefficients(t))
    coef = coef.cuda()
    lat_a = _r(2, 4, 16, 16, seed=1, dtype=torch.float32)
    lat_b = lat_a.clone()
    noise = _r(len(ts), 2 * 4 * 16 * 16, seed=2, dtype=torch.float32)
    step_idx = torch.zeros(1, device="cuda", dtype=torch.int32)
    t_dev = torch.full((1,), float(ts[0]), device="cuda")
    for i, t in enumerate(ts):
        mo = _r(4, 4, 16, 16, seed=10 + i, dtype=torch.float32)
        ops.cfg_ddpm_step(mo, lat_a, noise[i].view(2, 4, 16, 16).contiguous(), 2, 3.0, *sched.coefficients(t))
        assert t_dev.item() == float(t) and step_idx.item() == i
        ops.cfg_ddpm_step_table(mo, lat_b, noise, 2, 3.0, coef, step_idx)
        ops.advance_step(step_idx, coef, t_dev)
        assert torch.equal(lat_a, lat_b), f"step {i}"
    assert step_idx.item() == 0 and t_dev.item() == float(ts[0])  # wrapped: the loop can be replayed


def test_groupnorm_window_kernel_everywhere():
    """The window-major cluster kernel forced at every shape it fits (MVD_GN_ROWS=2, read once per process -> its own
    interpreter): small images, both channel-window alignments, two-source inputs."""
    import subprocess, sys, textwrap

    code = textwrap.dedent("""
        import sys, torch, torch.nn.functional as F
        sys.path.insert(0, %r)
        from mvd_b200 import ops
        g = torch.Generator(device="cuda").manual_seed(0)
        for (n, hw, c1, c2) in [(8, 64, 1280, 1280), (8, 256, 1280, 0), (8, 1024, 640, 0), (8, 1024, 640, 320), (4, 4096, 320, 0),
                                (16, 1024, 640, 0), (6, 1000, 320, 0), (8, 256, 640, 0), (5, 640, 64, 0)]:
            x1 = (torch.randn(n, hw, c1, device="cuda", generator=g) + 0.3).to(torch.bfloat16)
            x2 = (2 * torch.randn(n, hw, c2, device="cuda", generator=g)).to(torch.bfloat16) if c2 else None
            C = c1 + c2
            gm = (1 + 0.2 * torch.randn(C, device="cuda", generator=g)).to(torch.bfloat16)
            bt = (0.2 * torch.randn(C, device="cuda", generator=g)).to(torch.bfloat16)
            out = ops.groupnorm(x1, gm, bt, groups=32, eps=1e-5, silu=True, x2=x2)
            xin = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], -1)
            ref = F.silu(F.group_norm(xin.permute(0, 2, 1), 32, gm.float(), bt.float(), 1e-5).permute(0, 2, 1))
            err = (out.float() - ref).abs().max().item()
            assert err <= 2.5e-2 + 1e-2 * ref.abs().max().item(), (n, hw, c1, c2, err)
            again = ops.groupnorm(x1, gm, bt, groups=32, eps=1e-5, silu=True, x2=x2)
            assert torch.equal(out, again), "not deterministic"
        print("ok")
    """) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MVD_GN_ROWS="2")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
