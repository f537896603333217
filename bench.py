#!/usr/bin/env python
"""bench.py — MV denoise steps/s (SD2.1 + MVD adapter, 512^2) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # ours (CUDA kernels), 1 process per GPU
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port), host cores

One "step" = one denoise step of ONE object: 4 target views x CFG 2 = 8 UNet samples at 64x64 latents
(BASELINE.json configs[1]): CFG duplication, MultiViewUNet forward (camera FiLM + image conditioning + 32 adapter
processors), CFG combine + DDPM step. Reference-UNet features and their K/V are step-invariant and cached
(never counted). Weights are random-init SD2.1 architecture (865.9M + adapters), inputs synthetic (seeded).

  value  : steps/s with everything resident in HBM; one captured CUDA graph replayed K times, CUDA events.
  e2e    : the same step through the public DenoiseSession API with HOST (pinned) latents + variance noise copied
           in and the updated latents copied out inside the timed region, every step.
  N > 1  : `value` = STRONG scaling of ONE object (configs[2]): its 4 views x 2 CFG branches are sharded over the
           ranks — views at N <= 4, view x CFG branch at N = 8 —, reference features normalised over the FULL batch on
           every rank (attention.py:95-103 couples the batch); the only per-step exchange is the CFG pair's prediction
           at N = 8 (NCCL all_gather of 128 KB inside the captured step). "sample_parallel" additionally reports the
           zero-communication replica mode (configs[4]: one object per GPU, aggregate object-steps/s).
  parity : (N = 1) the timed model — same weights — run on 2 views x 32^2 latents and compared with the fp32 CPU
           oracle (north-star bound: normalised max-abs <= 2e-2, cosine >= 0.999), plus a finiteness check and checksum
           of the latents the timed configuration produced.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

VIEWS, LATENT, CFG = 4, 64, 2
WORKLOAD = ("configs[1]: one SD2.1 UNet + MV-adapter denoise step, 4 views at 512^2 (64^2 latent), CFG batch 2")
METRIC = "MV denoise steps/s (SD2.1+adapter 512^2, 4 views x CFG 2)"
GUIDANCE = 3.0
INFER_STEPS = 50
FLOPS_PER_STEP = 8.80e12  # 8 samples x 1151.6 GF minus the cached K/V projections
CROSS_VIEW = False


def select_workload(name: str):
    """c1 (default, the driver's line): BASELINE.json configs[1]/[2]. c3: configs[3] — 8 views at 768^2 (96^2 latents),
    camera + image conditioning, cross-view reference mode: every view attends over the reference tokens of ALL 8 views
    (S_kv = 73 728 at the top sites), no CFG; sharded one view per GPU at N = 8."""
    global VIEWS, LATENT, CFG, WORKLOAD, METRIC, GUIDANCE, FLOPS_PER_STEP, CROSS_VIEW
    if name == "c1":
        return
    if name != "c3":
        raise SystemExit(f"unknown workload {name}")
    VIEWS, LATENT, CFG, GUIDANCE, CROSS_VIEW = 8, 96, 1, 1.0, True
    WORKLOAD = ("configs[3]: one SD2.1 UNet + MV-adapter denoise step, 8 views at 768^2 (96^2 latent), cross-view "
                "reference attention over all 8 views' tokens (S_kv = 73728 at the top sites), camera + image conditioning")
    METRIC = "MV denoise steps/s (SD2.1+adapter 768^2, 8 views, cross-view K/V)"
    from flops import unet_flops

    f = unet_flops(LATENT, skv_factor=VIEWS)
    # per sample: everything except the step-invariant K/V projections of the reference tokens (half of ref_proj's
    # S_kv-sized terms: to_k_ref and to_v_ref; they are computed once per object)
    hw_terms = unet_flops(LATENT, skv_factor=1)["ref_proj"] / 2  # to_q_ref + to_out_ref
    FLOPS_PER_STEP = VIEWS * (f["base"] + f["ref_attn"] + hw_terms)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows)}


# ------------------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------------------
def build_pipeline(dev):
    import mvd_b200

    torch.manual_seed(0)
    pipe = mvd_b200.create_mvd_pipeline(None, dtype=torch.bfloat16, img_ref_scale=1.0, cam_modulation_strength=1.0,
                                        matched_batch_cfg=True, device=dev, cross_view_reference=CROSS_VIEW)
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():  # ref branch != original branch
        for n, p in pipe.unet.named_parameters():
            if any(s in n for s in ("to_k_ref", "to_v_ref", "to_out_ref.0.weight")):
                p.add_((torch.randn(p.shape, generator=g, device=dev) * 0.02).to(p.dtype))
    return pipe


def build_session(dev, views_local: int, view0: int, cfg_local: int, cfg_branch: int, use_graph=True, pipe=None,
                  sharded=None, emulated=False):
    """The slice of the object this rank owns: views [view0, view0+views_local), CFG branches
    (both if cfg_local == 2, else only `cfg_branch`: 0 = uncond, 1 = cond)."""
    from helpers import synthetic_inputs
    from mvd_b200.pipeline import DenoiseSession

    if pipe is None:
        pipe = build_pipeline(dev)
    inp = synthetic_inputs(VIEWS, LATENT, CFG)
    vs = slice(view0, view0 + views_local)
    text_u, text_c = inp["text"][:VIEWS][vs], inp["text"][VIEWS * (CFG - 1):][vs]
    unet = pipe.unet
    # reference features are computed over ALL views on every rank (step-invariant; normalisation statistics
    # couple the batch, attention.py:95-103), this rank's processors then use the rows of its own samples
    unet.shard = None
    if (views_local * cfg_local < VIEWS * CFG) if sharded is None else sharded:
        unet.shard = dict(view0=view0, views_local=views_local, views_total=VIEWS, cfg_total=CFG, cfg_branch=cfg_branch,
                          ie_text=inp["text"][VIEWS * (CFG - 1):].to(dev).contiguous(), emulated=bool(emulated))
    if cfg_local == 2:
        sess = DenoiseSession(pipe, text_c, INFER_STEPS, GUIDANCE, text_u, inp["source_camera"][vs],
                              inp["target_camera"][vs], inp["source_latents"], LATENT, use_cuda_graph=use_graph,
                              pos_proj=inp["pos_proj"])
    else:
        sess = DenoiseSession(pipe, text_c if (cfg_branch or CFG == 1) else text_u, INFER_STEPS, 1.0, None, inp["source_camera"][vs],
                              inp["target_camera"][vs], inp["source_latents"], LATENT, use_cuda_graph=use_graph,
                              pos_proj=inp["pos_proj"])
    noises = torch.stack([torch.randn(VIEWS, 4, LATENT, LATENT, generator=torch.Generator().manual_seed(6 + i))
                          for i in range(INFER_STEPS)])[:, vs]
    sess.reset(inp["latents"][vs], noises)
    return pipe, sess, inp, noises


def attention_roofline(dev, pk, how):
    """Dominant-kernel roofline: the cross-view / self attention core at the top site (B=8, h=5, S=4096, d=64),
    timed alone with CUDA events on inputs larger than L2 (see below)."""
    from mvd_b200 import ops

    B, H, S = VIEWS * CFG, 5, LATENT * LATENT
    C = H * 64
    g = torch.Generator(device=dev).manual_seed(3)
    s_kv = VIEWS * S if CROSS_VIEW else S
    # Inputs larger than L2 instead of a flush: N_SETS independent (q, k, v, out) sets used round-robin, so every launch
    # reads operands that the launches in between have evicted (126 MB L2), and LAUNCHES_PER_REGION launches share one
    # CUDA-event pair (no host gap inside the timed region).
    N_SETS, LAUNCHES_PER_REGION, REGIONS = 4, 12, 3
    sets = []
    for _ in range(N_SETS):
        qkv = torch.randn(B, S, 3 * C, device=dev, generator=g).to(torch.bfloat16)
        q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
        if CROSS_VIEW:  # the reference branch: one K/V sequence of all views' tokens shared by every sample
            kv = torch.randn(1, s_kv, 2 * C, device=dev, generator=g).to(torch.bfloat16)
            k, v = kv[:, :, :C].expand(B, -1, -1), kv[:, :, C:].expand(B, -1, -1)
        sets.append((q, k, v, torch.empty(B, S, C, device=dev, dtype=torch.bfloat16)))
    set_bytes = sum(t.numel() * 2 for t in (sets[0][0], sets[0][3])) + 2 * s_kv * C * 2 * (1 if CROSS_VIEW else B)

    def timed(**kw):
        for i in range(N_SETS):
            q, k, v, o = sets[i]
            ops.attention(q, k, v, H, out=o, **kw)
        times = []
        for _ in range(REGIONS):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(LAUNCHES_PER_REGION):
                q, k, v, o = sets[i % N_SETS]
                ops.attention(q, k, v, H, out=o, **kw)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / LAUNCHES_PER_REGION)
        return sum(times) / len(times)

    ms = timed()  # a stand-alone launch: the last partial wave is split along S_kv over the idle SMs
    # Inside the step: with persistent CTAs (default) the two branches of an adapter block follow each other on one
    # stream, each launched exactly as above. Without (MVD_ATTN_PERSIST=0) they run concurrently on two streams and
    # fill each other's last wave, so neither splits (co_units): the same kernel, timed alone in that configuration.
    persistent = ops.attention_is_persistent(B, H, S, s_kv, torch.cuda.get_device_properties(dev).multi_processor_count)
    ms_in_step_cfg = ms if persistent else timed(co_units=ops.attention_units(B, H, S))
    flops = 4.0 * S * s_kv * C * B
    achieved = flops / (ms * 1e-3) / 1e12
    peak = pk["bf16_tflops"]
    traffic = None
    if not CROSS_VIEW:  # the committed ncu capture is of the configs[1] shape
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "attn_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
    return {"kernel": f"{'attn_pair_persist_kernel' if persistent else 'attn_pair_kernel'} (B={B},h={H},Sq={S},Skv={s_kv},d=64)", "bound": "tensor", "achieved": round(achieved, 1),
            "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4), "traffic": traffic,
            "peak_source": f"{how} burst bf16 GEMM", "ms_per_launch": round(ms, 4),
            "timing": f"{LAUNCHES_PER_REGION} launches per CUDA-event pair, mean of {REGIONS}; {N_SETS} input sets used round-robin "
                      f"({N_SETS * set_bytes / 2**20:.0f} MiB > 126 MB L2) instead of an L2 flush",
            "flops_per_launch": flops,
            "as_launched_in_step": {"ms_per_launch": round(ms_in_step_cfg, 4),
                                    "achieved": round(flops / (ms_in_step_cfg * 1e-3) / 1e12, 1),
                                    "frac": round(flops / (ms_in_step_cfg * 1e-3) / 1e12 / peak, 4),
                                    "note": "persistent CTAs, launched as timed above" if persistent else
                                            "no KV-split tail: the concurrent sibling launch fills the last wave"}}


_ORACLE = None


def oracle_model(state_dict=None):
    """The fp32 CPU oracle of MultiViewUNet at full SD2.1 width (test infrastructure: used here only as the checker
    and as the CPU baseline). Built once; `state_dict` (the timed model's weights) is loaded when given."""
    global _ORACLE
    from oracle.mv_adapter import MultiViewUNetOracle

    if _ORACLE is None:
        torch.manual_seed(0)
        if state_dict is not None:
            try:  # skip the random init of 1.8 G parameters: build on the meta device, adopt the given tensors
                with torch.device("meta"):
                    m = MultiViewUNetOracle(None, img_ref_scale=1.0, cam_modulation_strength=1.0, matched_batch_cfg=True)
                m.load_state_dict({k: v.detach().float().cpu() for k, v in state_dict.items()}, assign=True)
                state_dict = None
            except Exception:
                m = MultiViewUNetOracle(None, img_ref_scale=1.0, cam_modulation_strength=1.0, matched_batch_cfg=True)
        else:
            m = MultiViewUNetOracle(None, img_ref_scale=1.0, cam_modulation_strength=1.0, matched_batch_cfg=True)
        _ORACLE = m.eval()
    if state_dict is not None:
        _ORACLE.load_state_dict({k: v.detach().float().cpu() for k, v in state_dict.items()})
    return _ORACLE


SAMPLE_VIEWS = 1  # the CPU arm times 1 of the 4 views (both CFG branches, full 64^2 latents, full-width model)


def cpu_sample_step(m, inp):
    """One bounded sample of the configs[1] step on the host: SAMPLE_VIEWS view(s) x CFG 2 at 64^2 latents through the
    full-width oracle, INCLUDING the frozen reference-UNet re-run the reference performs every step (mvd_unet.py:287)."""
    x = torch.cat([inp["latents"][:SAMPLE_VIEWS]] * CFG)
    text = torch.cat([inp["text"][:VIEWS][:SAMPLE_VIEWS], inp["text"][VIEWS:][:SAMPLE_VIEWS]])
    with torch.no_grad():
        return m(x, 981, text, inp["source_camera"][:SAMPLE_VIEWS], inp["target_camera"][:SAMPLE_VIEWS],
                 inp["source_latents"][:SAMPLE_VIEWS], pos_proj=inp["pos_proj"]).sample


def cpu_baseline(reps=2, warmup=1, model=None):
    """CPU baseline of the same metric: the oracle port of the reference's per-step work on the host cores (all
    threads), timed on a bounded sample of configs[1] — SAMPLE_VIEWS of the 4 views at the real latent size and CFG —
    and scaled by the view count (the step is batch-linear: every view runs the same layers on the same shapes)."""
    from flops import unet_flops
    from helpers import synthetic_inputs

    torch.set_num_threads(os.cpu_count() or 1)
    m = model if model is not None else oracle_model()
    inp = synthetic_inputs(VIEWS, LATENT, CFG)
    times = []
    for _ in range(warmup + reps):
        t0 = time.time()
        cpu_sample_step(m, inp)
        times.append(time.time() - t0)
    timed = times[warmup:]
    sec = sum(timed) / len(timed)
    f = unet_flops(LATENT)
    sample_flops = SAMPLE_VIEWS * (CFG * f["total"] + f["base"])  # the reference re-runs the frozen UNet every step
    steps_per_s = 1.0 / (sec * VIEWS / SAMPLE_VIEWS)
    return {"value": steps_per_s, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle port (fp32 torch CPU, full SD2.1 width) of MultiViewUNet.forward incl. the per-step "
                      f"frozen-UNet re-run: {SAMPLE_VIEWS} of the {VIEWS} views x CFG {CFG} at {LATENT}x{LATENT} latents = "
                      f"{sample_flops / 1e12:.2f} TFLOP in {sec:.2f} s (mean of {reps} after {warmup} warm-up), x"
                      f"{VIEWS // SAMPLE_VIEWS} views for the full step", "cpu_tflops": sample_flops / sec / 1e12,
            "sample_seconds": sec, "sample_fraction": SAMPLE_VIEWS / VIEWS}, timed


def processor_c0(dev):
    """configs[0] of BASELINE.json — the reference's own CPU-runnable case: ONE ImageCrossAttentionProcessor forward
    (original self-attention + reference branch, attention.py:48-188) on hidden [4,4096,320], reference [4,320,64,64],
    5 heads. CPU: the oracle's processor (checked to 0.0 max-abs against the live reference processor by
    oracle/gen_golden.py), fp32, all host threads, 2 warm-ups, best of 5. GPU: the product's processor through the
    same call (K/V cache warm, as in every step but the first), CUDA events, 3 warm-ups, mean of 20. Parity between
    the two is printed."""
    from helpers import metrics
    from oracle import mv_adapter
    from oracle.sd21_unet import Attention as OAttention

    import mvd_b200
    from mvd_b200 import unet as punet

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(11)
    name, c, heads, views, side = "down_block_0_attn_0_self", 320, 5, 4, 64
    torch.manual_seed(11)
    attn_o = OAttention(c, heads, 64)
    proc_o = mv_adapter.make_processor(name, attn_o, img_ref_scale=1.0)
    with torch.no_grad():
        for w in (proc_o.to_k_ref.weight, proc_o.to_v_ref.weight, proc_o.to_out_ref[0].weight):
            w.add_(0.02 * torch.randn(w.shape, generator=g))
    attn_o.processor = proc_o
    hidden = torch.randn(views, side * side, c, generator=g)
    ref = torch.randn(views, c, side, side, generator=g)
    times = []
    with torch.no_grad():
        for _ in range(7):
            t0 = time.time()
            y_o = attn_o(hidden, ref_hidden_states={name: ref})
            times.append(time.time() - t0)
    cpu_ms = min(times[2:]) * 1e3
    flops = views * (2 * 4 * side ** 4 * c + 8 * 2 * side * side * c * c)  # 2 SDPA + 8 projections per view
    attn_p = punet.Attention(c, heads, 64)
    attn_p.load_state_dict(attn_o.state_dict(), strict=False)
    attn_p = attn_p.to(dev, torch.bfloat16)
    proc_p = mvd_b200.get_attention_processor_for_module(name, attn_p, img_ref_scale=1.0)
    proc_p.load_state_dict(proc_o.state_dict())
    attn_p.processor = proc_p.to(dev, torch.bfloat16)
    hs, rf = hidden.to(dev, torch.bfloat16), {name: ref.to(dev)}
    with torch.no_grad():
        for _ in range(3):
            y_p = attn_p(hs, ref_hidden_states=rf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            y_p = attn_p(hs, ref_hidden_states=rf)
        e1.record()
        torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / 20
    m = metrics(y_p.float().cpu(), y_o)
    return {"workload": "configs[0]: one ImageCrossAttentionProcessor forward, hidden [4,4096,320], ref [4,320,64,64], "
                        "5 heads", "cpu_ms": round(cpu_ms, 2), "cpu_tflops": round(flops / cpu_ms / 1e9, 3),
            "cpu": "oracle processor (0.0 max-abs vs the live reference processor), fp32, best of 5",
            "cores": torch.get_num_threads(), "gpu_ms": round(gpu_ms, 4),
            "gpu": "product processor, bf16, eager launches (no graph), reference K/V cached, mean of 20",
            "normalised_max_abs": round(m["rel"], 5), "cosine": round(m["cos"], 6)}


def parity_check(pipe, sess_latents, dev):
    """The timed model against the oracle at a size the oracle finishes in seconds (2 views x 32^2, cfg 1), same
    weights; and sanity of what the timed configuration itself produced."""
    from helpers import metrics, synthetic_inputs

    unet = pipe.unet
    saved_shard = unet.shard
    unet.shard = None
    o = oracle_model(unet.state_dict())
    inp = synthetic_inputs(2, 32, 1)
    cam = unet.camera_encoder
    saved_proj = cam._pos_proj
    cam.set_positional_projection(inp["pos_proj"])
    with torch.no_grad():
        y_p = unet(inp["latents"].to(dev), 981, inp["text"].to(dev), inp["source_camera"].to(dev),
                   inp["target_camera"].to(dev), inp["source_latents"].to(dev)).sample
        y_o = o(inp["latents"], 981, inp["text"], inp["source_camera"], inp["target_camera"], inp["source_latents"],
                pos_proj=inp["pos_proj"]).sample
    cam._pos_proj = saved_proj
    unet.shard = saved_shard
    m = metrics(y_p, y_o)
    lat = sess_latents.detach().float().cpu()
    return {"vs": "fp32 CPU oracle, same weights, 2 views x 32x32 latents, full SD2.1 width", "max_abs": round(m["max_abs"], 5),
            "normalised_max_abs": round(m["rel"], 5), "cosine": round(m["cos"], 6),
            "ok": bool(m["rel"] <= 2e-2 and m["cos"] >= 0.999),
            "timed_config_latents": {"finite": bool(torch.isfinite(lat).all()), "abs_mean": round(float(lat.abs().mean()), 6),
                                     "checksum": round(float(lat.double().sum()), 4)}}


def build_rank_session(dev, sp, pipe, use_graph, guidance=GUIDANCE):
    """Session of the slice of the object plan `sp` gives this rank (+ the CFG pair exchange when the pair is split)."""
    from mvd_b200 import dist as mdist

    _, s2, inp, noises = build_session(dev, sp["views_local"], sp["view0"], sp["cfg_local"], sp["cfg_branch"],
                                       use_graph=use_graph, pipe=pipe, sharded=sp["world"] > 1,
                                       emulated=sp.get("emulated", False))
    if sp["cfg_local"] == 1 and CFG == 2:
        if sp.get("emulated"):  # no partner rank: the pair's other prediction is a copy of ours (same kernels, no NCCL)
            from mvd_b200 import ops
            gathered = torch.empty((CFG,) + tuple(s2.latents.shape), device=dev, dtype=torch.float32)

            def step(sess=s2):
                out = sess.forward_unet(sess.latents)
                gathered[0].copy_(out)
                gathered[1].copy_(out)
                ops.cfg_ddpm_step_table(gathered, sess.latents, sess.noise_table, CFG, guidance, sess.coef, sess.step_idx)
                sess.advance()

            s2._eager_step = step
            s2.guidance = guidance
        else:
            mdist.install_cfg_pair_exchange(s2, sp, guidance)
    return s2, inp, noises


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
    if world not in (1, 2, 4, 8):
        raise SystemExit("supported GPU counts: 1, 2, 4, 8")
    from mvd_b200 import dist as mdist

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # Headline = ONE object (BASELINE.json configs[1] at N = 1, configs[2] at N > 1): its V x cfg samples are split over
    # the ranks (strong scaling). Every rank holds the full model; reference features cover all views on every rank.
    sp = mdist.shard_plan(VIEWS, CFG, world, rank)
    if args.shard_of > 1:  # single-GPU stand-in for rank 0 of an N-rank job (profiling the per-rank step)
        if world != 1:
            raise SystemExit("--shard-of emulates one rank on ONE GPU")
        sp = mdist.shard_plan(VIEWS, CFG, args.shard_of, 0)
        sp["emulated"] = True
    pipe = build_pipeline(dev)
    if args.profile:  # one eager step between cudaProfilerStart/Stop (ncu --profile-from-start off)
        sess, inp, noises = build_rank_session(dev, sp, pipe, use_graph=False)
        for _ in range(2):
            sess.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        sess.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled_launches": sess.launches_per_step}))
        return

    finished = threading.Event()
    emitted = threading.Lock()
    result = {}

    def emit(extra=None):
        if rank != 0 or not emitted.acquire(blocking=False):  # exactly one JSON line
            return
        line = dict(result.get("line") or {"metric": METRIC, "value": None, "unit": "steps/s", "n_gpus": world,
                                            "error": "the timed section did not complete"})
        line.update(extra or {})
        print(json.dumps(line))
        sys.stdout.flush()

    if world > 1:
        def watchdog():  # a hung collective must never hang the bench: report what we have and leave
            if not finished.wait(420.0):
                emit({"watchdog": "timed out after 420 s"})
                os._exit(0)

        threading.Thread(target=watchdog, daemon=True).start()

    sess, graph_used, err = None, True, None
    for use_graph in (True, False):  # NCCL inside a captured step (N = 8) falls back to eager launches if capture fails
        try:
            sess, inp, noises = build_rank_session(dev, sp, pipe, use_graph)
            if use_graph:
                sess.capture()
            torch.cuda.synchronize()
            graph_used = use_graph
            break
        except Exception as exc:  # noqa: BLE001
            err = f"{type(exc).__name__}: {exc}"[:300]
            sess = None
    if sess is None:
        raise SystemExit(f"bench.py: could not build the sharded session: {err}")

    for _ in range(max(args.warmup, 3)):
        sess.step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        sess.step()
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    timed_latents = sess.latents.clone()

    # ---- e2e: host (pinned) latents + noise in, updated latents out, every step, through the session API
    vs = slice(sp["view0"], sp["view0"] + sp["views_local"])
    lat_host = inp["latents"][vs].contiguous().pin_memory()
    noise_host = noises[0].contiguous().pin_memory()
    out_host = torch.empty_like(lat_host).pin_memory()
    noise_slot = sess.noise_table[0].view_as(sess.latents)
    k_e2e = max(3, min(args.steps, 20))

    def e2e_step():
        sess.latents.copy_(lat_host, non_blocking=True)
        noise_slot.copy_(noise_host, non_blocking=True)
        sess.step()
        out_host.copy_(sess.latents, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        e2e_step()
    barrier()
    e2e_sps = k_e2e / max_over_ranks(time.perf_counter() - t0)

    pk, how = peaks()
    steps_per_s = args.steps / (ms_total * 1e-3)  # steps of the ONE object per second
    result["line"] = {
        "metric": METRIC, "value": round(steps_per_s, 3), "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD + ", bf16, random-init weights", "views": VIEWS, "cfg": CFG, "latent": LATENT,
                   "parallelism": sp["desc"], "samples_per_gpu": sp["views_local"] * sp["cfg_local"],
                   "l2": "working set (1.9 GB weights + activations) >> 126 MB L2; no flush",
                   "cuda_graph": graph_used, "step_invariant_cached": "reference-UNet features, reference/text K/V, camera emb"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_sps, 3), "unit": "steps/s", "h2d_bytes_per_step": lat_host.numel() * 4 * 2 * world,
                "d2h_bytes_per_step": out_host.numel() * 4 * world, "steps": k_e2e},
        "gpu_launches": int(sess.launches_per_step * args.steps * world),
        "launches_per_step": int(sess.launches_per_step),
        "step_tflops": round(FLOPS_PER_STEP * steps_per_s / 1e12, 1),
        "step_frac_of_sustained_peak": round(FLOPS_PER_STEP * steps_per_s / 1e12 /
                                             (max(world, args.shard_of) * pk["bf16_tflops_sustained"]), 4),
    }

    # ---- N > 1: the zero-communication replica mode (configs[4]) as a side figure
    if world > 1 and not args.no_replicas:
        try:
            pipe.unet.shard = None
            del sess
            rp = mdist.shard_plan(VIEWS, CFG, 1, 0)
            _, s2, _, _ = build_session(dev, rp["views_local"], rp["view0"], rp["cfg_local"], rp["cfg_branch"],
                                        use_graph=True, pipe=pipe)
            s2.capture()
            for _ in range(3):
                s2.step()
            barrier()
            k_sp = max(3, min(args.steps, 10))
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            for _ in range(k_sp):
                s2.step()
            v1.record()
            barrier()
            sp_ms = max_over_ranks(v0.elapsed_time(v1))
            result["line"]["sample_parallel"] = {
                "value": round(world * k_sp / (sp_ms * 1e-3), 3), "unit": "object-steps/s over all GPUs", "scaling": "weak",
                "ms_per_step": round(sp_ms / k_sp, 3), "steps": k_sp,
                "parallelism": f"one object (4 views x CFG 2) per GPU x{world}, no data-path collective (configs[4])"}
        except Exception as exc:  # noqa: BLE001
            result["line"]["sample_parallel"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if args.shard_of > 1:
        result["line"]["emulated_rank_of"] = args.shard_of
    if world == 1 and rank == 0 and args.shard_of == 1:
        result["line"]["roofline"] = attention_roofline(dev, pk, how)
        if not args.no_parity and args.workload == "c1":
            try:
                result["line"]["parity"] = parity_check(pipe, timed_latents, dev)
            except Exception as exc:  # noqa: BLE001
                result["line"]["parity"] = {"error": f"{type(exc).__name__}: {exc}"[:300], "ok": False}
        if not args.no_cpu_baseline and args.workload == "c1":
            result["line"]["cpu_baseline"] = cpu_baseline()[0]
            try:
                result["line"]["cpu_baseline"]["c0"] = processor_c0(dev)
            except Exception as exc:  # noqa: BLE001
                result["line"]["cpu_baseline"]["c0"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    finished.set()
    emit()
    # All measurements are done and reported. Skip the NCCL teardown on purpose: destroy_process_group() after a
    # CUDA graph that captured NCCL work has been observed to hang on this stack; leaving through os._exit is safe
    # because nothing is left to flush but stdout.
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


# ------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle port), all host threads
# ------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """K timed steps (after W warm-ups), each step = the bounded sample cpu_sample_step() of the configs[1] step (1 of
    the 4 views, both CFG branches, 64^2 latents, full-width model, incl. the per-step frozen-UNet re-run). `value` is
    the full-configuration figure (x4 views); `ms_per_step` is the measured time of one timed (sample) step, so that
    steps x ms_per_step is the wall time of the timed region."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k, w = max(1, args.steps), max(0, args.warmup)
    best, timed = cpu_baseline(reps=k, warmup=w)
    v = best["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": k, "warmup": w,
        "ms_per_step": 1e3 * sum(timed) / len(timed), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + ", fp32 on the host cores, random-init weights; every timed step is a bounded "
                               f"sample ({SAMPLE_VIEWS} of {VIEWS} views, see cpu_baseline.sample) and `value` is scaled to "
                               "the full step", "views": VIEWS, "cfg": CFG, "latent": LATENT,
                   "sample_fraction": SAMPLE_VIEWS / VIEWS},
        "cpu_baseline": best,
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c1", choices=["c1", "c3"],
                    help="c1 = BASELINE configs[1]/[2] (default, the driver's line); c3 = configs[3]: 8 views at 768^2, cross-view K/V")
    ap.add_argument("--shard-of", type=int, default=1, help="single GPU: run rank 0's share of an N-rank view-sharded job "
                    "(no NCCL; for profiling the per-rank step)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed model (N = 1)")
    ap.add_argument("--no-replicas", action="store_true", help="N > 1: skip the sample-parallel (replica) side figure")
    ap.add_argument("--profile", action="store_true", help="run one eager step inside cudaProfilerStart/Stop and exit")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
