"""CPU tests of the multi-GPU host logic: the shard plan, and (gloo, world size 2) the per-step CFG-pair exchange
that the view x CFG sharded layout performs with NCCL on GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvd_b200.dist import local_sample_index, shard_plan


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_plan_partitions_all_samples(world):
    views, cfg = 4, 2
    seen = []
    for rank in range(world):
        p = shard_plan(views, cfg, world, rank)
        idx = local_sample_index(views, cfg, p["view0"], p["views_local"], p["cfg_local"], p["cfg_branch"])
        assert len(idx) == views * cfg // world
        seen += idx
        if world <= views:
            assert p["cfg_local"] == cfg and p["pair"] is None  # both CFG branches local: no exchange
            # local order is [uncond(views), cond(views)] of the rank's own views
            assert idx == [b * views + p["view0"] + v for b in range(cfg) for v in range(p["views_local"])]
        else:
            assert p["cfg_local"] == 1 and rank in p["pair"] and len(p["pair"]) == cfg
            assert p["cfg_branch"] == rank % cfg and p["view0"] == rank // cfg
    assert sorted(seen) == list(range(views * cfg))


@pytest.mark.parametrize("views,cfg,world", [(4, 2, 1), (4, 2, 2), (4, 2, 4), (4, 2, 8), (1, 2, 2), (8, 1, 8), (8, 2, 16)])
def test_local_sample_index_stays_inside_the_reference_batch(views, cfg, world):
    """Indices address the [cfg_total * V] matched-batch reference features; every rank's index set is in range and
    the union covers each sample exactly once (the out-of-range gather ADVICE.md flagged cannot be built)."""
    seen = []
    for rank in range(world):
        p = shard_plan(views, cfg, world, rank)
        idx = local_sample_index(views, cfg, p["view0"], p["views_local"], p["cfg_local"], p["cfg_branch"])
        assert all(0 <= i < views * cfg for i in idx)
        seen += idx
    assert sorted(seen) == list(range(views * cfg))


def test_shard_plan_rejects_bad_world_sizes():
    with pytest.raises(ValueError):
        shard_plan(4, 2, 3, 0)
    with pytest.raises(ValueError):
        shard_plan(4, 2, 16, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _pair_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.noise_schedule import DDPMOracle

        plan = shard_plan(1, 2, world, rank)  # one view, CFG 2 -> one branch per rank
        g = torch.Generator().manual_seed(0)
        lat = torch.randn(1, 4, 8, 8, generator=g)
        preds = torch.randn(2, 1, 4, 8, 8, generator=g)  # [branch, ...] what each rank's UNet would output
        noise = torch.randn(1, 4, 8, 8, generator=g)
        mine = preds[plan["cfg_branch"]].contiguous()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)  # the exchange install_cfg_pair_exchange() does with NCCL
        u, c = gathered[0], gathered[1]
        sched = DDPMOracle()
        sched.set_timesteps(50)
        out = sched.step(u + 3.0 * (c - u), 981, lat, noise)
        ref = sched.step(preds[0] + 3.0 * (preds[1] - preds[0]), 981, lat, noise)
        ret[rank] = float((out - ref).abs().max())
    finally:
        dist.destroy_process_group()


def test_cfg_pair_exchange_gloo_world2():
    """Both ranks of a view end the step with identical, correct latents (bit-exact with the single-process step)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_pair_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == 0.0 and ret[1] == 0.0
