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
  N > 1  : `value` = sample-parallel weak scaling (configs[4]): every GPU denoises its own object, no data-path
           collective; aggregate object-steps/s. "view_sharded" additionally reports strong scaling of ONE object
           (configs[2]): views (N <= 4) or view x CFG branch (N = 8) sharded over ranks, reference features
           normalised over the FULL batch on every rank; only per-step exchange = the CFG pair's prediction at N = 8
           (NCCL all_gather of 128 KB).
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
GUIDANCE = 3.0
INFER_STEPS = 50
FLOPS_PER_STEP = 8.80e12  # 8 samples x 1151.6 GF minus the cached K/V projections


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
                                        matched_batch_cfg=True, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():  # ref branch != original branch
        for n, p in pipe.unet.named_parameters():
            if any(s in n for s in ("to_k_ref", "to_v_ref", "to_out_ref.0.weight")):
                p.add_((torch.randn(p.shape, generator=g, device=dev) * 0.02).to(p.dtype))
    return pipe


def build_session(dev, views_local: int, view0: int, cfg_local: int, cfg_branch: int, use_graph=True, pipe=None):
    """The slice of the object this rank owns: views [view0, view0+views_local), CFG branches
    (both if cfg_local == 2, else only `cfg_branch`: 0 = uncond, 1 = cond)."""
    from helpers import synthetic_inputs
    from mvd_b200.pipeline import DenoiseSession

    if pipe is None:
        pipe = build_pipeline(dev)
    inp = synthetic_inputs(VIEWS, LATENT, CFG)
    vs = slice(view0, view0 + views_local)
    text_u, text_c = inp["text"][:VIEWS][vs], inp["text"][VIEWS:][vs]
    unet = pipe.unet
    # reference features are computed over ALL views on every rank (step-invariant; normalisation statistics
    # couple the batch, attention.py:95-103), this rank's processors then use the rows of its own samples
    unet.shard = None
    if views_local < VIEWS:
        unet.shard = dict(view0=view0, views_local=views_local, views_total=VIEWS, cfg_total=CFG, cfg_branch=cfg_branch,
                          ie_text=inp["text"][VIEWS:].to(dev).contiguous())
    if cfg_local == 2:
        sess = DenoiseSession(pipe, text_c, INFER_STEPS, GUIDANCE, text_u, inp["source_camera"][vs],
                              inp["target_camera"][vs], inp["source_latents"], LATENT, use_cuda_graph=use_graph,
                              pos_proj=inp["pos_proj"])
    else:
        sess = DenoiseSession(pipe, text_c if cfg_branch else text_u, INFER_STEPS, 1.0, None, inp["source_camera"][vs],
                              inp["target_camera"][vs], inp["source_latents"], LATENT, use_cuda_graph=use_graph,
                              pos_proj=inp["pos_proj"])
    noises = torch.stack([torch.randn(VIEWS, 4, LATENT, LATENT, generator=torch.Generator().manual_seed(6 + i))
                          for i in range(INFER_STEPS)])[:, vs]
    sess.reset(inp["latents"][vs], noises)
    return pipe, sess, inp, noises


def attention_roofline(dev, pk, how):
    """Dominant-kernel roofline: the cross-view / self attention core at the top site (B=8, h=5, S=4096, d=64),
    timed alone with CUDA events, L2 flushed between iterations."""
    from mvd_b200 import ops

    B, H, S = VIEWS * CFG, 5, LATENT * LATENT
    C = H * 64
    g = torch.Generator(device=dev).manual_seed(3)
    qkv = torch.randn(B, S, 3 * C, device=dev, generator=g).to(torch.bfloat16)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    out = torch.empty(B, S, C, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for _ in range(3):
        ops.attention(q, k, v, H, out=out)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.attention(q, k, v, H, out=out)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    flops = 4.0 * S * S * C * B
    achieved = flops / (ms * 1e-3) / 1e12
    peak = pk["bf16_tflops"]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "attn_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    return {"kernel": "attn_fwd2_kernel (B=8,h=5,Sq=Skv=4096,d=64)", "bound": "tensor", "achieved": round(achieved, 1),
            "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4), "traffic": traffic,
            "peak_source": f"{how} burst bf16 GEMM", "ms_per_launch": round(ms, 4),
            "flops_per_launch": flops}


def cpu_baseline(sample_views=2, latent=32, reps=2):
    """Bounded CPU sample of the same path (oracle port of the reference's per-step work): full-width SD2.1 UNet +
    adapters INCLUDING the frozen reference-UNet re-run the reference performs every step (mvd_unet.py:287), on
    `sample_views` views at `latent`^2 latents, fp32, all host threads; scaled by algorithmic FLOPs to the
    configs[1] step (4 views x CFG 2 at 64^2)."""
    from flops import unet_flops
    from oracle.mv_adapter import MultiViewUNetOracle
    from helpers import synthetic_inputs

    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    m = MultiViewUNetOracle(None, img_ref_scale=1.0, cam_modulation_strength=1.0).eval()
    inp = synthetic_inputs(sample_views, latent, 1)
    args = (inp["latents"], 981, inp["text"], inp["source_camera"], inp["target_camera"], inp["source_latents"])
    times = []
    with torch.no_grad():
        for _ in range(1 + reps):  # first call is the warm-up
            t0 = time.time()
            m(*args, pos_proj=inp["pos_proj"])
            times.append(time.time() - t0)
    sec = min(times[1:])
    f = unet_flops(latent)
    sample_flops = sample_views * (f["total"] + f["base"])  # the reference re-runs the frozen UNet every step
    full_flops = VIEWS * CFG * unet_flops(LATENT)["total"] + VIEWS * unet_flops(LATENT)["base"]
    steps_per_s = (sample_flops / sec) / full_flops
    return {"value": steps_per_s, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle port (fp32 torch CPU) of MultiViewUNet.forward incl. the per-step frozen-UNet re-run, "
                      f"{sample_views} views x {latent}x{latent} latents = {sample_flops / 1e9:.0f} GFLOP in {sec:.2f} s "
                      f"(best of {reps} after 1 warm-up), scaled by FLOPs to the {full_flops / 1e12:.2f} TFLOP step the "
                      f"reference executes", "cpu_tflops": sample_flops / sec / 1e12}


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
    # default multi-GPU mode = sample-parallel (BASELINE.json configs[4]): every GPU denoises its OWN
    # object (4 views x CFG 2), no data-path collective -> weak scaling; the view-sharded strong-scaling number of
    # ONE object (configs[2]) is measured afterwards and reported under "view_sharded".
    plan = mdist.shard_plan(VIEWS, CFG, 1, 0)
    plan["desc"] = "single GPU" if world == 1 else \
        f"sample-parallel x{world}: one object (4 views x CFG 2) per GPU, no data-path collective"
    pipe, sess, inp, noises = build_session(dev, plan["views_local"], plan["view0"], plan["cfg_local"], plan["cfg_branch"],
                                            use_graph=not args.profile)
    if args.profile:  # one eager step between cudaProfilerStart/Stop (ncu --profile-from-start off)
        for _ in range(2):
            sess.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        sess.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled_launches": sess.launches_per_step}))
        return
    sess.capture()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

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
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---- e2e: host (pinned) latents + noise in, updated latents out, every step, through the session API
    lat_host = inp["latents"][plan["view0"]:plan["view0"] + plan["views_local"]].contiguous().pin_memory()
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
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_sps = world * k_e2e / float(e2e_s.item())

    emitted = threading.Lock()

    def emit(view_sharded):
        if not emitted.acquire(blocking=False):  # exactly one JSON line
            return
        pk, how = peaks()
        steps_per_s = world * args.steps / (ms_total * 1e-3)  # object-steps per second over all GPUs
        line = {
            "metric": "MV denoise steps/s (SD2.1+adapter 512^2, 4 views x CFG 2)", "value": round(steps_per_s, 3),
            "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD + ", bf16, random-init weights", "views": VIEWS, "cfg": CFG, "latent": LATENT,
                       "parallelism": plan["desc"], "l2": "working set (1.9 GB weights + activations) >> 126 MB L2; no flush",
                       "cuda_graph": True, "step_invariant_cached": "reference-UNet features, reference/text K/V, camera emb"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_sps, 3), "unit": "steps/s", "h2d_bytes_per_step": lat_host.numel() * 4 * 2,
                    "d2h_bytes_per_step": out_host.numel() * 4, "steps": k_e2e},
            "gpu_launches": int(sess.launches_per_step * args.steps * world),
            "launches_per_step": int(sess.launches_per_step),
            "step_tflops": round(FLOPS_PER_STEP * steps_per_s / 1e12, 1),
            "step_frac_of_sustained_peak": round(FLOPS_PER_STEP * steps_per_s / 1e12 / (world * pk["bf16_tflops_sustained"]), 4),
        }
        if view_sharded is not None:
            line["view_sharded"] = view_sharded
        if world == 1:
            line["roofline"] = attention_roofline(dev, pk, how)
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
        sys.stdout.flush()

    # ---- strong scaling of ONE object (configs[2]): views / CFG branches sharded over the ranks
    view_sharded = None
    finished = threading.Event()
    if world > 1:
        import torch.distributed as dist

        def watchdog():  # the optional section must never hang the bench: after 120 s report what we have and leave
            if not finished.wait(120.0):
                if rank == 0:
                    emit({"error": "timed out after 120 s", "parallelism": "view-sharded"})
                os._exit(0)

        threading.Thread(target=watchdog, daemon=True).start()
    run_strong = world > 1 and (world <= VIEWS or args.strong8)
    if world > 1 and not run_strong:
        view_sharded = {"skipped": "N = V*cfg needs a per-step NCCL exchange of the CFG pair (mvd_b200/dist.py); "
                                   "run with --strong8 to time it"}
    if run_strong:
        sp = mdist.shard_plan(VIEWS, CFG, world, rank)
        k_vs = max(3, min(args.steps, 10))
        for use_graph in (True, False):  # NCCL inside a captured step (N = 8) falls back to eager launches if needed
            try:
                _, s2, _, _ = build_session(dev, sp["views_local"], sp["view0"], sp["cfg_local"], sp["cfg_branch"],
                                            use_graph=use_graph, pipe=pipe)
                if sp["cfg_local"] == 1 and CFG == 2:
                    mdist.install_cfg_pair_exchange(s2, sp, GUIDANCE)
                if use_graph:
                    s2.capture()
                for _ in range(3):
                    s2.step()
                barrier()
                v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                v0.record()
                for _ in range(k_vs):
                    s2.step()
                v1.record()
                barrier()
                vms = torch.tensor([v0.elapsed_time(v1)], device=dev)
                dist.all_reduce(vms, op=dist.ReduceOp.MAX)
                view_sharded = {"value": round(k_vs / (float(vms.item()) * 1e-3), 3), "unit": "steps/s of one object",
                                "ms_per_step": round(float(vms.item()) / k_vs, 3), "scaling": "strong",
                                "parallelism": sp["desc"], "cuda_graph": use_graph, "steps": k_vs}
                break
            except Exception as exc:  # noqa: BLE001
                view_sharded = {"error": f"{type(exc).__name__}: {exc}"[:300], "cuda_graph": use_graph}
        pipe.unet.shard = None
    finished.set()
    if rank == 0:
        emit(view_sharded)
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k, w = max(1, min(args.steps, 3)), 1
    best = cpu_baseline(reps=k)
    v = best["value"]
    line = {
        "impl": "reference", "metric": "MV denoise steps/s (SD2.1+adapter 512^2, 4 views x CFG 2)", "value": v,
        "unit": "steps/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": k, "warmup": w,
        "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD + ", fp32 on the host cores, random-init weights; timed on a bounded sample "
                               "(cpu_baseline.sample) and scaled by FLOPs", "views": VIEWS, "cfg": CFG, "latent": LATENT},
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
    ap.add_argument("--strong8", action="store_true", help="also time the view x CFG sharded mode at 8 GPUs")
    ap.add_argument("--profile", action="store_true", help="run one eager step inside cudaProfilerStart/Stop and exit")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
