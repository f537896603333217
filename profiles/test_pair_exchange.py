"""2-rank check of the view x CFG sharded step (mvd_b200/dist.py install_cfg_pair_exchange): 1 view, CFG 2,
one CFG branch per rank, NCCL all_gather of the pair's prediction; eager and CUDA-graph; compared with the
single-process CFG step. Run: torchrun --nproc-per-node 2 profiles/test_pair_exchange.py"""
import os, sys, signal
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
signal.alarm(150)  # never hang the box
import torch
import torch.distributed as dist
import mvd_b200
from mvd_b200 import dist as mdist
from mvd_b200.pipeline import DenoiseSession
from mvd_b200.unet import tiny_config
from helpers import synthetic_inputs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
torch.manual_seed(0)
m = mvd_b200.MultiViewUNet(tiny_config(), dtype=torch.bfloat16, img_ref_scale=1.0, cam_modulation_strength=1.0,
                           matched_batch_cfg=True).to(dev, dtype=torch.bfloat16).eval()
V, L, steps, G = 1, 16, 4, 3.0
inp = synthetic_inputs(V, L, cfg=2, text_dim=64)
m.camera_encoder.set_positional_projection(inp["pos_proj"])
noises = torch.stack([torch.randn(V, 4, L, L, generator=torch.Generator().manual_seed(6 + i)) for i in range(steps)])
sched = mvd_b200.ShiftSNRScheduler.from_scheduler(mvd_b200.DDPMScheduler(), shift_mode="interpolated", shift_scale=6.0,
                                                  scheduler_class=mvd_b200.DDPMScheduler)
pipe = mvd_b200.MVDPipeline(unet=m, scheduler=sched)
# reference: both branches on this GPU
ref = DenoiseSession(pipe, inp["text"][V:].to(dev), steps, G, inp["text"][:V].to(dev), inp["source_camera"],
                     inp["target_camera"], inp["source_latents"].to(dev), L, use_cuda_graph=False)
ref.reset(inp["latents"].to(dev), noises)
ref.run()
plan = mdist.shard_plan(V, 2, world, rank)
for graph in (False, True):
    m.shard = dict(view0=0, views_local=1, views_total=1, cfg_total=2, cfg_branch=plan["cfg_branch"],
                   ie_text=inp["text"][V:].to(dev).contiguous())
    text = inp["text"][V:] if plan["cfg_branch"] else inp["text"][:V]
    s = DenoiseSession(pipe, text.to(dev), steps, 1.0, None, inp["source_camera"], inp["target_camera"],
                       inp["source_latents"].to(dev), L, use_cuda_graph=graph)
    mdist.install_cfg_pair_exchange(s, plan, G)
    s.reset(inp["latents"].to(dev), noises)
    s.run()
    torch.cuda.synchronize()
    err = (s.latents - ref.latents).abs().max().item()
    print(f"rank {rank} graph={graph}: max |sharded - single| = {err:.3e}", flush=True)
    cos = torch.nn.functional.cosine_similarity(s.latents.flatten(), ref.latents.flatten(), dim=0).item()
    print(f"rank {rank} graph={graph}: cosine {cos:.6f}", flush=True)
    assert cos > 0.999  # 4 CFG-3 steps amplify bf16-level tile-order differences; direction must agree
    m.shard = None
print(f"rank {rank} OK", flush=True)
os._exit(0)  # NCCL teardown after a captured collective hangs on this stack; nothing left to flush
