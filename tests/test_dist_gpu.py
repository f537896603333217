"""2-GPU check of the view x CFG sharded denoise step (mvd_b200/dist.py install_cfg_pair_exchange; BASELINE.json
configs[2] at N = V*cfg): one view, CFG 2, one CFG branch per rank, NCCL all_gather of the pair's prediction before the
fused CFG + DDPM kernel; eager and CUDA-graph replay, compared with the single-GPU CFG session on the same inputs.
Needs two visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`); skipped otherwise."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    import signal

    signal.alarm(240)  # never hang the box
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import mvd_b200
    from helpers import metrics, synthetic_inputs
    from mvd_b200 import dist as mdist
    from mvd_b200.pipeline import DenoiseSession
    from mvd_b200.unet import tiny_config

    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    torch.manual_seed(0)
    m = mvd_b200.MultiViewUNet(tiny_config(), dtype=torch.bfloat16, img_ref_scale=1.0, cam_modulation_strength=1.0,
                               matched_batch_cfg=True).to(dev, dtype=torch.bfloat16).eval()
    V, L, steps, G = 1, 16, 4, 3.0
    inp = synthetic_inputs(V, L, cfg=2, text_dim=64)
    m.camera_encoder.set_positional_projection(inp["pos_proj"])
    noises = torch.stack([torch.randn(V, 4, L, L, generator=torch.Generator().manual_seed(6 + i)) for i in range(steps)])
    sched = mvd_b200.ShiftSNRScheduler.from_scheduler(mvd_b200.DDPMScheduler(), shift_mode="interpolated",
                                                      shift_scale=6.0, scheduler_class=mvd_b200.DDPMScheduler)
    pipe = mvd_b200.MVDPipeline(unet=m, scheduler=sched)
    # single-GPU session: both CFG branches on this GPU
    ref = DenoiseSession(pipe, inp["text"][V:].to(dev), steps, G, inp["text"][:V].to(dev), inp["source_camera"],
                         inp["target_camera"], inp["source_latents"].to(dev), L, use_cuda_graph=False)
    ref.reset(inp["latents"].to(dev), noises)
    ref.run(1)
    ref_step1 = ref.latents.clone()
    ref.run(steps - 1)
    plan = mdist.shard_plan(V, 2, world, rank)
    worst = 0.0
    for graph in (False, True):
        m.shard = dict(view0=0, views_local=1, views_total=1, cfg_total=2, cfg_branch=plan["cfg_branch"],
                       ie_text=inp["text"][V:].to(dev).contiguous())
        text = inp["text"][V:] if plan["cfg_branch"] else inp["text"][:V]
        s = DenoiseSession(pipe, text.to(dev), steps, 1.0, None, inp["source_camera"], inp["target_camera"],
                           inp["source_latents"].to(dev), L, use_cuda_graph=graph)
        mdist.install_cfg_pair_exchange(s, plan, G)
        s.reset(inp["latents"].to(dev), noises)
        s.run(1)
        torch.cuda.synchronize()
        m1 = metrics(s.latents, ref_step1)  # ONE step: the north-star bound applies
        s.run(steps - 1)
        torch.cuda.synchronize()
        mm = metrics(s.latents, ref.latents)  # the roll-out: differences of one step are fed back (CFG 3) and compound
        worst = max(worst, mm["rel"])
        ret[(rank, graph)] = (m1["rel"], m1["cos"], mm["rel"], mm["cos"])
        m.shard = None
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)  # NCCL teardown after a captured collective has been seen to hang on this stack


def _xview_worker(rank, world, port, ret):
    import signal

    signal.alarm(240)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import mvd_b200
    from helpers import metrics, synthetic_inputs
    from mvd_b200 import dist as mdist
    from mvd_b200.unet import tiny_config

    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    torch.manual_seed(0)
    m = mvd_b200.MultiViewUNet(tiny_config(), dtype=torch.bfloat16, img_ref_scale=1.0, cam_modulation_strength=1.0,
                               matched_batch_cfg=True, cross_view_reference=True).to(dev, dtype=torch.bfloat16).eval()
    V, L = 2, 16
    inp = synthetic_inputs(V, L, cfg=1, text_dim=64)
    m.camera_encoder.set_positional_projection(inp["pos_proj"])
    x, text, src = inp["latents"].to(dev), inp["text"].to(dev), inp["source_latents"].to(dev)
    with torch.no_grad():
        full = m(x, 621, text, inp["source_camera"].to(dev), inp["target_camera"].to(dev), src).sample
        p = mdist.shard_plan(V, 1, world, rank)
        vs = slice(p["view0"], p["view0"] + p["views_local"])
        # every rank encodes only its own view; the per-site reference tokens travel by NCCL all_gather
        m.shard = dict(view0=p["view0"], views_local=p["views_local"], views_total=V, cfg_total=1, cfg_branch=0,
                       ie_text=text.contiguous())
        y = m(x[vs].contiguous(), 621, text[vs].contiguous(), inp["source_camera"][vs].to(dev),
              inp["target_camera"][vs].to(dev), src).sample
    torch.cuda.synchronize()
    mm = metrics(y, full[vs])
    ret[rank] = (mm["rel"], mm["cos"])
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_cross_view_reference_all_gather_matches_single_gpu():
    """BASELINE configs[3] / north-star collective: views split over ranks, the step-invariant per-view reference
    tokens of all 16 sites all-gathered once (NCCL over NVLink); each rank's prediction equals the rows of the
    single-GPU cross-view forward."""
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.spawn(_xview_worker, args=(2, _free_port(), ret), nprocs=2, join=False)
    for p in ctx.processes:
        p.join(300)
    got = dict(ret)
    assert len(got) == 2, f"ranks did not finish: {got}"
    for rank, (rel, cos) in got.items():
        print(f"rank {rank}: normalised max-abs {rel:.3e} cos {cos:.6f}")
        assert rel <= 2e-2 and cos >= 0.9999, (rank, rel, cos)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_cfg_pair_sharded_step_matches_single_gpu():
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    ctx = mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=False)
    for p in ctx.processes:
        p.join(300)
    got = dict(ret)
    assert len(got) == 4, f"ranks did not finish: {got}"
    for key, (rel1, cos1, rel, cos) in got.items():
        print(f"rank {key[0]} graph={key[1]}: one step normalised max-abs {rel1:.3e} cos {cos1:.6f}; "
              f"{4}-step roll-out {rel:.3e} cos {cos:.6f}")
        # same samples, same kernels, batch 1 instead of 2: tile decomposition (stream-K, tile width) and hence bf16
        # rounding differ. One step must hold the north-star bound; over the roll-out the differences are fed back
        # through CFG 3 and compound, so only the cosine bound and a looser max-abs are asserted there.
        assert rel1 <= 2e-2 and cos1 >= 0.999, (key, rel1, cos1)
        assert rel <= 5e-2 and cos >= 0.999, (key, rel, cos)
