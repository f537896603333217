import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import mvd_b200
from mvd_b200.pipeline import DenoiseSession
from mvd_b200.unet import tiny_config
from helpers import synthetic_inputs
torch.manual_seed(0)
m = mvd_b200.MultiViewUNet(tiny_config(), dtype=torch.bfloat16, img_ref_scale=1.0, cam_modulation_strength=1.0, matched_batch_cfg=True).to("cuda", dtype=torch.bfloat16).eval()
V, L, steps = 2, 16, 6
inp = synthetic_inputs(V, L, cfg=2, text_dim=64)
m.camera_encoder.set_positional_projection(inp["pos_proj"])
noises = torch.stack([torch.randn(V, 4, L, L, generator=torch.Generator().manual_seed(6 + i)) for i in range(steps)])
sched = mvd_b200.ShiftSNRScheduler.from_scheduler(mvd_b200.DDPMScheduler(), shift_mode="interpolated", shift_scale=6.0, scheduler_class=mvd_b200.DDPMScheduler)
pipe = mvd_b200.MVDPipeline(unet=m, scheduler=sched)
def mk(graph):
    s = DenoiseSession(pipe, inp["text"][V:].cuda(), steps, 3.0, inp["text"][:V].cuda(), inp["source_camera"], inp["target_camera"], inp["source_latents"].cuda(), L, use_cuda_graph=graph)
    s.reset(inp["latents"].cuda(), noises)
    return s
a, b = mk(False), mk(True)
for i in range(steps):
    a.step(); b.step(); torch.cuda.synchronize()
    print(i, "eager-session vs graph-session", (a.latents - b.latents).abs().max().item(), "step_idx", a.step_idx.item(), b.step_idx.item(), "t", a.t_dev.item(), b.t_dev.item())
c = mk(False)
for i in range(steps):
    c.step()
print("eager-session rerun vs eager-session", (a.latents - c.latents).abs().max().item())
kw = dict(prompt_embeds=inp["text"][V:].cuda(), negative_prompt_embeds=inp["text"][:V].cuda(), latents=inp["latents"].cuda(), num_inference_steps=steps, guidance_scale=3.0,
          source_camera=inp["source_camera"], target_camera=inp["target_camera"], source_image_latents=inp["source_latents"].cuda(), variance_noises=noises, height=L * 8, width=L * 8)
e = pipe(**kw)["latents"]
print("pipeline eager vs eager-session", (e - a.latents).abs().max().item())
# first-step forensics
s = mk(False)
inp2 = torch.cat([s.latents] * 2)
with torch.no_grad():
    o1 = m(sample=inp2, timestep=s.t_dev, encoder_hidden_states=s.text, **s.extra).sample.clone()
    o2 = m(sample=inp2, timestep=831, encoder_hidden_states=s.text, **s.extra).sample.clone()
    txt = torch.cat([inp["text"][:V].cuda(), inp["text"][V:].cuda()])
    o3 = m(sample=inp2, timestep=831, encoder_hidden_states=txt, source_camera=inp["source_camera"].cuda(), target_camera=inp["target_camera"].cuda(), source_image_latents=inp["source_latents"].cuda()).sample.clone()
    o4 = m(sample=inp2, timestep=831, encoder_hidden_states=txt, cross_attention_kwargs={}, source_camera=inp["source_camera"].cuda(), target_camera=inp["target_camera"].cuda(), source_image_latents=inp["source_latents"].cuda()).sample.clone()
print("tensor-t vs int-t", (o1 - o2).abs().max().item(), " new tensors", (o2 - o3).abs().max().item(), "with kwargs {}", (o3 - o4).abs().max().item())
print("text equal", torch.equal(txt, s.text), s.text.dtype, txt.dtype)
