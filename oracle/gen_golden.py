"""Generate tests/golden/*.npz from the REFERENCE ITSELF (authoring container only).

Run:  python -m oracle.gen_golden            (needs /root/reference; never runs on the GPU box)

What it does
  1. imports the reference's own src/models/attention.py, src/models/camera_encoder.py and
     src/training/scheduler.py unmodified from /root/reference (with a 3-line stub for the `icecream`
     dependency, which is only used for printing),
  2. runs them on small seeded inputs,
  3. checks that the oracle restatement (oracle/mv_adapter.py, oracle/noise_schedule.py) reproduces the
     reference outputs to fp32 round-off — this is what pins the oracle,
  4. stores inputs' seeds + the reference outputs as small fixtures that travel to the GPU box.

The module weights are regenerated from seeds at test time (torch's CPU RNG is deterministic for a given
torch build; the fixtures also store a weight checksum so a silent RNG change is detected, not mis-read as
a parity failure).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def _import_reference():
    if "icecream" not in sys.modules:
        stub = types.ModuleType("icecream")
        stub.ic = lambda *a, **k: None
        sys.modules["icecream"] = stub
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.models.attention as ref_attention  # noqa: E402
    import src.models.camera_encoder as ref_camera  # noqa: E402
    import src.training.scheduler as ref_sched  # noqa: E402

    return ref_attention, ref_camera, ref_sched


# ---- shared, seeded case builders (also imported by tests/) ---------------------------------------------
PROC_CASES = {
    # name: (C, heads, B_q, HW_side, ref kind, B_ref, cross)
    "self_4d": dict(C=128, heads=2, bq=2, side=8, kind="4d", bref=2, cross=False),
    "cross_4d": dict(C=128, heads=2, bq=2, side=8, kind="4d", bref=2, cross=True),
    "self_3d_multiview": dict(C=128, heads=2, bq=2, side=8, kind="3d", bref=2, cross=False, skv=4 * 64),
    "self_cfg_literal": dict(C=64, heads=1, bq=2, side=8, kind="4d", bref=1, cross=False),
    "self_c320": dict(C=320, heads=5, bq=1, side=8, kind="4d", bref=1, cross=False),
}


def build_proc_case(name: str):
    """Returns (attn_module, hidden, text, ref) for a case; everything drawn from fixed seeds."""
    from .sd21_unet import Attention

    c = PROC_CASES[name]
    torch.manual_seed(1000 + sorted(PROC_CASES).index(name))
    attn = Attention(c["C"], c["heads"], 64, cross_attention_dim=128 if c["cross"] else None)
    g = torch.Generator().manual_seed(7)
    hidden = torch.randn(c["bq"], c["side"] ** 2, c["C"], generator=g)
    text = torch.randn(c["bq"], 11, 128, generator=g) if c["cross"] else None
    if c["kind"] == "4d":
        ref = torch.randn(c["bref"], c["C"], c["side"], c["side"], generator=g) * 1.5 + 0.3
    else:
        ref = torch.randn(c["bref"], c["skv"], c["C"], generator=g) * 1.5 + 0.3
    return attn, hidden, text, ref


def perturb_processor(proc, seed: int = 1):
    """make the ref branch differ from the original one."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in (proc.to_k_ref.weight, proc.to_v_ref.weight, proc.to_out_ref[0].weight):
            p.add_(torch.randn(p.shape, generator=g) * 0.02)


def build_camera_case(V: int = 4):
    from .mv_adapter import CameraEncoderOracle

    torch.manual_seed(55)
    dims = {"down_0": 64, "up_0": 128, "mid": 128, "output": 4}
    enc = CameraEncoderOracle(output_dim=1024, hidden_dim=512, modulation_hidden_dims=dims, modulation_strength=0.8)
    g = torch.Generator().manual_seed(9)
    src = torch.randn(V, 3, 4, generator=g)
    tgt = torch.randn(V, 3, 4, generator=g)
    proj = torch.randn(1024, 1020, generator=g) / np.sqrt(1020)
    x = torch.randn(V, 64, 5, 5, generator=g)
    return enc, src, tgt, proj, x, dims


def _checksum(module) -> float:
    return float(sum(p.detach().double().abs().sum() for p in module.parameters()))


def main():
    ref_attention, ref_camera, ref_sched = _import_reference()
    from . import mv_adapter, noise_schedule

    os.makedirs(GOLDEN, exist_ok=True)
    out = {}

    # ---- 1. processor -----------------------------------------------------------------------------------
    for name in PROC_CASES:
        attn, hidden, text, ref = build_proc_case(name)
        rp = ref_attention.get_attention_processor_for_module(name, attn, img_ref_scale=0.7)
        op = mv_adapter.make_processor(name, attn, img_ref_scale=0.7)
        for (kn, a), (_, b) in zip(rp.state_dict().items(), op.state_dict().items()):
            assert torch.equal(a, b), f"{name}: weight seeding differs at {kn}"
        perturb_processor(rp)
        perturb_processor(op)
        with torch.no_grad():
            y_ref = rp(attn, hidden, encoder_hidden_states=text, ref_hidden_states={name: ref})
            y_missing = rp(attn, hidden, encoder_hidden_states=text, ref_hidden_states={"other": ref})
            y_orc = op(attn, hidden, encoder_hidden_states=text, ref_hidden_states={name: ref})
            y_orc_missing = op(attn, hidden, encoder_hidden_states=text, ref_hidden_states=None)
        err = (y_ref - y_orc).abs().max().item()
        assert err < 1e-5, f"{name}: oracle differs from the reference processor by {err}"
        assert torch.equal(y_missing, y_orc_missing), f"{name}: fall-through differs"
        out[f"proc/{name}/y"] = y_ref.numpy()
        out[f"proc/{name}/y_noref"] = y_missing.numpy()
        out[f"proc/{name}/wsum"] = np.float64(_checksum(rp) + _checksum(attn))
        print(f"processor {name}: reference vs oracle max abs {err:.2e}")

    # ---- 2. camera encoder ------------------------------------------------------------------------------
    enc, src, tgt, proj, x, dims = build_camera_case()
    rc = ref_camera.CameraEncoder(output_dim=1024, hidden_dim=512, modulation_hidden_dims=dims, modulation_strength=0.8)
    rc.load_state_dict(enc.state_dict())
    real_randn = torch.randn

    def fake_randn(*shape, **kw):  # the reference draws the projection on every call (camera_encoder.py:153)
        assert tuple(shape) == (1024, 1020)
        return proj * np.sqrt(1020)

    torch.randn = fake_randn
    try:
        with torch.no_grad():
            e_ref = rc.encode_cameras(src, tgt)
    finally:
        torch.randn = real_randn
    with torch.no_grad():
        e_orc = enc.encode_cameras(src, tgt, proj)
        f_ref = rc.apply_modulation(x, "down_0", e_ref)
        f_orc = enc.film(x, "down_0", e_orc)
        f_tuple = rc.apply_modulation((x, x), "down_0", e_ref)
        f_unknown = rc.apply_modulation(x, "mid_0", e_ref)  # the name the reference's mid hook uses
    err = (e_ref - e_orc).abs().max().item()
    assert err < 1e-5, f"camera embedding differs by {err}"
    assert (f_ref - f_orc).abs().max().item() < 1e-5
    assert torch.equal(f_tuple[1], x) and torch.equal(f_unknown, x)
    out["camera/embedding"] = e_ref.numpy()
    out["camera/film_down_0"] = f_ref.numpy()
    out["camera/wsum"] = np.float64(_checksum(rc))
    print(f"camera encoder: reference vs oracle max abs {err:.2e}; 'mid_0' hook name is a no-op: True")

    # ---- 3. schedule ------------------------------------------------------------------------------------
    class _Sched:  # the two attributes scheduler.py reads
        def __init__(self, betas):
            self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
            self.config = types.SimpleNamespace(num_train_timesteps=betas.shape[0])

    base = noise_schedule.sd21_betas()
    captured = {}

    class _Cls:
        @staticmethod
        def from_config(config, trained_betas=None):
            captured["betas"] = torch.tensor(trained_betas)
            return captured

    ref_sched.ShiftSNRScheduler.from_scheduler(_Sched(base), shift_mode="interpolated", shift_scale=6.0,
                                               scheduler_class=_Cls)
    b_ref = captured["betas"]
    b_orc = noise_schedule.shifted_betas(base, 6.0, "interpolated")
    err = (b_ref - b_orc).abs().max().item()
    assert err < 1e-7, f"shifted betas differ by {err}"
    ref_sched.ShiftSNRScheduler.from_scheduler(_Sched(base), shift_mode="default", shift_scale=6.0, scheduler_class=_Cls)
    assert (captured["betas"] - noise_schedule.shifted_betas(base, 6.0, "default")).abs().max().item() < 1e-7
    out["sched/shifted_betas"] = b_ref.numpy()
    print(f"shifted betas: reference vs oracle max abs {err:.2e}")

    ref_sched.ShiftSNRScheduler.from_scheduler(_Sched(base), shift_mode="default", shift_scale=6.0, scheduler_class=_Cls)
    out["sched/shifted_betas_default"] = captured["betas"].numpy()
    t_all = torch.arange(base.shape[0])
    snr_ref = ref_sched.compute_snr(t_all, _Sched(base))
    assert (snr_ref - noise_schedule.snr_from_betas(base)).abs().max().item() <= 1e-6 * snr_ref.abs().max().item()
    out["sched/snr"] = snr_ref.numpy()
    rt = ref_sched.SNR_to_betas(snr_ref)
    assert (torch.as_tensor(rt) - base).abs().max().item() < 1e-6  # SNR_to_betas inverts compute_snr
    out["sched/betas_roundtrip"] = torch.as_tensor(rt).numpy()

    # ---- 4. camera encoder, piece by piece --------------------------------------------------------------
    with torch.no_grad():
        rel_ref = rc.compute_relative_transform(src, tgt)
        rel_orc = enc.relative_transform(src, tgt)
    for key, b in zip(("R", "T"), rel_orc):  # camera_encoder.py:107-120 returns {"R": ..., "T": ...}
        assert (rel_ref[key] - b).abs().max().item() < 1e-6, f"relative transform {key} differs"
        out[f"camera/relative_{key}"] = rel_ref[key].numpy()
    gx = torch.Generator().manual_seed(21)
    for mod_name, ch in dims.items():
        xm = torch.randn(src.shape[0], ch, 3, 4, generator=gx) * 1.7 - 0.2
        with torch.no_grad():
            a = rc.apply_modulation(xm, mod_name, e_ref)
            b = enc.film(xm, mod_name, e_orc)
        assert (a - b).abs().max().item() < 1e-5, f"FiLM at {mod_name} differs"
        out[f"camera/film_all/{mod_name}"] = a.numpy()

    # ---- 5. the look-at matrices the benchmark's synthetic cameras use (src/utils.py:51-85) --------------
    import src.utils as ref_utils  # noqa: E402
    import math

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import camera_matrix  # noqa: E402

    ring = []
    for V_ in (4, 8):
        for i in range(V_):
            az, el, r = math.radians(360.0 * i / V_), math.radians(20.0), 1.8
            pos = [r * math.cos(el) * math.sin(az), r * math.sin(el), r * math.cos(el) * math.cos(az)]
            m_ref = ref_utils.create_camera_matrix(pos, [0.0, 0.0, 0.0])
            assert (m_ref - camera_matrix(360.0 * i / V_)).abs().max().item() < 1e-6
            ring.append(m_ref.numpy())
    out["utils/camera_ring"] = np.stack(ring)
    print("camera pieces, FiLM at every modulator, look-at ring, SNR round trip: pinned")

    old_path = os.path.join(GOLDEN, "reference_adapter.npz")
    if os.path.exists(old_path):  # regenerating must not move any vector that is already committed
        prev = np.load(old_path)
        for k in prev.files:
            assert k in out and np.array_equal(prev[k], out[k]), f"golden vector {k} changed"
    np.savez_compressed(os.path.join(GOLDEN, "reference_adapter.npz"), **out)
    print("wrote", os.path.join(GOLDEN, "reference_adapter.npz"),
          os.path.getsize(os.path.join(GOLDEN, "reference_adapter.npz")), "bytes")


if __name__ == "__main__":
    main()
