"""Multi-GPU partitioning of one object's denoise step — one process per GPU, torch.distributed.

The V x cfg samples of a step are independent through the whole UNet once the reference features exist; the only
cross-sample coupling is the reference-feature normalisation (reference attention.py:95-103: statistics over the
batch). So every rank computes the (step-invariant, cached) reference features for ALL views, normalises over the
full batch exactly as a single GPU would, and then attends with the K/V rows of its own samples only.

  N <= V      : views are split; each rank keeps both CFG branches of its views -> no per-step communication.
  N == V*cfg  : one sample per rank; the two ranks of a view exchange their predictions (NCCL all_gather of one
                [1,4,L,L] fp32 tensor) before the fused CFG + DDPM kernel; both keep identical latents.
No collective is needed for K/V in the reference's semantics (K/V come from the frozen reference UNet, not from
the other views' live hidden states).
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import ops


def shard_plan(views: int, cfg: int, world: int, rank: int) -> Dict:
    if world <= views:
        if views % world:
            raise ValueError(f"{views} views do not split over {world} ranks")
        vl = views // world
        return dict(views_local=vl, view0=rank * vl, cfg_local=cfg, cfg_branch=0, pair=None, world=world, rank=rank,
                    desc=f"view-sharded x{world} ({vl} views x cfg {cfg} per GPU), no per-step collective"
                    if world > 1 else "single GPU")
    if world != views * cfg:
        raise ValueError(f"world size {world} must be <= {views} or == {views * cfg}")
    return dict(views_local=1, view0=rank // cfg, cfg_local=1, cfg_branch=rank % cfg, world=world, rank=rank,
                pair=[(rank // cfg) * cfg + i for i in range(cfg)],
                desc=f"view x CFG sharded x{world} (1 sample per GPU), per-step NCCL all_gather of the CFG pair")


def local_sample_index(views_total: int, cfg_total: int, view0: int, views_local: int, cfg_local: int,
                       cfg_branch: int) -> List[int]:
    """Global sample indices (order [uncond views..., cond views...]) of a rank's local batch."""
    branches = range(cfg_total) if cfg_local == cfg_total else [cfg_branch]
    return [b * views_total + view0 + v for b in branches for v in range(views_local)]


def install_cfg_pair_exchange(sess, plan: Dict, guidance: float):
    """N == V*cfg: replace the session's step by forward -> all_gather(pair) -> fused CFG + DDPM."""
    import torch.distributed as dist

    world, cfg = plan["world"], len(plan["pair"])
    groups = {}
    for base in range(0, world, cfg):  # every rank creates every group (collective call)
        ranks = list(range(base, base + cfg))
        groups[base] = dist.new_group(ranks)
    group = groups[plan["pair"][0]]
    gathered = torch.empty((cfg,) + tuple(sess.latents.shape), device=sess.latents.device, dtype=torch.float32)

    def step():
        out = sess.forward_unet(sess.latents)
        dist.all_gather_into_tensor(gathered, out.contiguous(), group=group)
        ops.cfg_ddpm_step_table(gathered, sess.latents, sess.noise_table, cfg, guidance, sess.coef, sess.step_idx)
        sess.advance()

    sess._eager_step = step
    sess.guidance = guidance
