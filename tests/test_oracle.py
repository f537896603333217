"""CPU tests: the oracle restatement against the golden vectors produced by the reference itself
(oracle/gen_golden.py), plus the structural known answers of SURVEY.md 8(c)."""
import os

import numpy as np
import pytest
import torch

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_adapter.npz"))


@pytest.mark.parametrize("case", ["self_4d", "cross_4d", "self_3d_multiview", "self_cfg_literal", "self_c320"])
def test_oracle_processor_reproduces_reference(case):
    from oracle import mv_adapter
    from oracle.gen_golden import _checksum, build_proc_case, perturb_processor

    attn, hidden, text, ref = build_proc_case(case)
    proc = mv_adapter.make_processor(case, attn, img_ref_scale=0.7)
    perturb_processor(proc)
    assert abs(_checksum(proc) + _checksum(attn) - float(GOLDEN[f"proc/{case}/wsum"])) < 1e-6, \
        "seeded weights differ from the generator's (torch RNG changed?)"
    with torch.no_grad():
        y = proc(attn, hidden, encoder_hidden_states=text, ref_hidden_states={case: ref})
        y0 = proc(attn, hidden, encoder_hidden_states=text, ref_hidden_states=None)
    assert np.abs(y.numpy() - GOLDEN[f"proc/{case}/y"]).max() < 1e-5
    assert np.abs(y0.numpy() - GOLDEN[f"proc/{case}/y_noref"]).max() < 1e-5
    assert np.abs(y.numpy() - y0.numpy()).max() > 1e-2  # the reference branch really contributes


def test_oracle_camera_encoder_reproduces_reference():
    from oracle.gen_golden import _checksum, build_camera_case

    enc, src, tgt, proj, x, dims = build_camera_case()
    assert abs(_checksum(enc) - float(GOLDEN["camera/wsum"])) < 1e-6
    with torch.no_grad():
        e = enc.encode_cameras(src, tgt, proj)
        f = enc.film(x, "down_0", e)
        assert torch.equal(enc.film(x, "mid_0", e), x)  # the reference's mid hook name is not a modulator
    assert np.abs(e.numpy() - GOLDEN["camera/embedding"]).max() < 1e-5
    assert np.abs(f.numpy() - GOLDEN["camera/film_down_0"]).max() < 1e-5


def test_oracle_shifted_betas_reproduce_reference_and_kats():
    from oracle import noise_schedule as ns

    base = ns.sd21_betas()
    b = ns.shifted_betas(base, 6.0, "interpolated")
    assert np.abs(b.numpy() - GOLDEN["sched/shifted_betas"]).max() < 1e-7
    # SURVEY.md 8(c) KATs: alpha_bar monotone, snr'(0) = snr(0), snr'(999) = snr(999) / 6
    abar = torch.cumprod(1 - b, 0)
    assert bool((abar[1:] < abar[:-1]).all())
    snr, snr_s = ns.snr_from_betas(base), ns.snr_from_betas(b)
    assert abs(snr_s[0] / snr[0] - 1) < 1e-4 and abs(snr_s[-1] / (snr[-1] / 6) - 1) < 2e-3


def test_oracle_ddpm_step_identities():
    from oracle.noise_schedule import DDPMOracle

    s = DDPMOracle()
    ts = s.set_timesteps(50).tolist()
    assert ts[:3] == [981, 961, 941] and ts[-1] == 1 and len(ts) == 50
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    t = 501
    a = s.alphas_cumprod[t]
    xt = a ** 0.5 * x0 + (1 - a) ** 0.5 * eps
    v = a ** 0.5 * eps - (1 - a) ** 0.5 * x0  # exact v-prediction
    sa, sb, c0, ct, sg = s.coefficients(t)
    assert torch.allclose(sa * xt - sb * v, x0, atol=1e-5)  # x0 is recovered exactly
    prev = s.step(v, t, xt, None)
    a_prev = s.alphas_cumprod[t - 20]
    # posterior mean with the true x0 and eps equals the DDIM-style point on the next noise level's mean path
    assert torch.allclose(prev, c0 * x0 + ct * xt, atol=1e-5)
    assert abs(c0 * float(a ** 0.5) + ct - float(a_prev ** 0.5) / float(a ** 0.5) * float(a ** 0.5)) < 1.0  # finite
    assert s.coefficients(1)[4] > 0 and s.coefficients(0)[4] == 0


def test_unet_parameter_count_kat():
    """SD2.1 UNet = 865,910,724 parameters; adapter processors 410,560 / 1,640,320 / 6,557,440; CameraEncoder
    19,062,536 — for the oracle and for the product module tree."""
    from oracle.mv_adapter import CameraEncoderOracle, RefAttnProcessor
    from oracle.sd21_unet import UNet2DConditionModel as OU
    from mvd_b200.unet import UNet2DConditionModel as PU
    from mvd_b200 import CameraEncoder, ImageCrossAttentionProcessor

    with torch.device("meta"):
        o, p = OU(), PU()
        assert sum(q.numel() for q in o.parameters()) == 865_910_724
        assert sum(q.numel() for q in p.parameters()) == 865_910_724
        assert list(o.state_dict().keys()) == list(p.state_dict().keys())
        for c, h, n in ((320, 5, 410_560), (640, 10, 1_640_320), (1280, 20, 6_557_440)):
            assert sum(q.numel() for q in RefAttnProcessor("x", c, h).parameters()) == n
            assert sum(q.numel() for q in ImageCrossAttentionProcessor("x", c, h).parameters()) == n
        dims = {"down_0": 320, "down_1": 640, "down_2": 1280, "down_3": 1280, "up_0": 1280, "up_1": 1280, "up_2": 640,
                "up_3": 320, "mid": 1280, "output": 4}
        assert sum(q.numel() for q in CameraEncoderOracle(1024, 512, modulation_hidden_dims=dims).parameters()) == 19_062_536
        assert sum(q.numel() for q in CameraEncoder(1024, 512, modulation_hidden_dims=dims).parameters()) == 19_062_536


def test_flop_totals_kat():
    """804.3 GFLOP per sample for the base UNet at L=64 (conv 418.4 + linear 259.8 +
    self-attn 122.5 + text-attn 3.6) and 347.3 GFLOP for the adapter — counted from the oracle's module tree."""
    from flops import unet_flops

    f = unet_flops(64)
    assert abs(f["conv"] / 1e9 - 418.4) < 0.5 and abs(f["linear"] / 1e9 - 259.8) < 0.5
    assert abs(f["self_attn"] / 1e9 - 122.5) < 0.2 and abs(f["text_attn"] / 1e9 - 3.6) < 0.1
    assert abs(f["base"] / 1e9 - 804.3) < 1.0
    assert abs(f["ref_attn"] / 1e9 - 245.0) < 0.5 and abs(f["ref_proj"] / 1e9 - 102.3) < 0.5
    assert abs(f["total"] / 1e9 - 1151.6) < 1.5
    f96 = unet_flops(96)
    assert abs(f96["total"] / 1e9 - 3619.6) < 4.0


def test_oracle_tiny_unet_runs_and_names_match_reference_walk():
    from oracle.mv_adapter import MultiViewUNetOracle, feature_taps
    from oracle.sd21_unet import tiny_config

    torch.manual_seed(0)
    m = MultiViewUNetOracle(tiny_config())
    names = [n for n, _ in feature_taps(m.base_unet)]
    assert names == ["down_block_0_attn_0", "down_block_0_attn_1", "down_block_1_attn_0", "down_block_1_attn_1",
                     "down_block_2_attn_0", "down_block_2_attn_1", "mid_block_attn_0", "up_block_1_attn_0",
                     "up_block_1_attn_1", "up_block_1_attn_2", "up_block_2_attn_0", "up_block_2_attn_1",
                     "up_block_2_attn_2", "up_block_3_attn_0", "up_block_3_attn_1", "up_block_3_attn_2"]
    assert len(m.attention_layer_map) == 32
    keys = m.state_dict().keys()
    assert "base_unet.down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor.to_out_ref.0.bias" in keys
    assert "image_encoder.unet.conv_in.weight" in keys and "camera_encoder.modulators.output.3.bias" in keys


def test_oracle_camera_pieces_reproduce_reference():
    """Relative pose (camera_encoder.py:107-120) and the FiLM of EVERY modulator (camera_encoder.py:198-255), on the
    vectors the live reference produced (oracle/gen_golden.py section 4)."""
    from oracle.gen_golden import build_camera_case

    enc, src, tgt, proj, _, dims = build_camera_case()
    with torch.no_grad():
        r, t = enc.relative_transform(src, tgt)
        e = enc.encode_cameras(src, tgt, proj)
    assert np.abs(r.numpy() - GOLDEN["camera/relative_R"]).max() < 1e-6
    assert np.abs(t.numpy() - GOLDEN["camera/relative_T"]).max() < 1e-6
    gx = torch.Generator().manual_seed(21)
    for name, ch in dims.items():
        x = torch.randn(src.shape[0], ch, 3, 4, generator=gx) * 1.7 - 0.2
        with torch.no_grad():
            y = enc.film(x, name, e)
        assert np.abs(y.numpy() - GOLDEN[f"camera/film_all/{name}"]).max() < 1e-5, name
        assert (y - x).abs().max() > 1e-2


def test_synthetic_camera_ring_is_the_references_look_at():
    """tests/helpers.camera_matrix (the benchmark's synthetic cameras) == src/utils.py:51-85 create_camera_matrix."""
    from helpers import camera_matrix

    ring = [camera_matrix(360.0 * i / v) for v in (4, 8) for i in range(v)]
    assert np.abs(torch.stack(ring).numpy() - GOLDEN["utils/camera_ring"]).max() < 1e-6


def test_oracle_snr_and_default_shift_reproduce_reference():
    from oracle import noise_schedule as ns

    base = ns.sd21_betas()
    snr = ns.snr_from_betas(base)
    assert np.abs(snr.numpy() - GOLDEN["sched/snr"]).max() <= 1e-6 * float(GOLDEN["sched/snr"].max())
    assert np.abs(ns.betas_from_snr(snr).numpy() - GOLDEN["sched/betas_roundtrip"]).max() < 1e-6
    b = ns.shifted_betas(base, 6.0, "default")
    assert np.abs(b.numpy() - GOLDEN["sched/shifted_betas_default"]).max() < 1e-7
