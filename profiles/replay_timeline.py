"""Kernel timeline of the GRAPH-REPLAYED step (CUPTI through torch.profiler): true per-kernel durations and the idle
gaps between them, which ncu's serialised, low-clock launch list cannot show.
usage: python profiles/replay_timeline.py [shard_of]   (1 = the whole object on one GPU, 8 = one sample per GPU)"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from mvd_b200 import dist as mdist

shard_of = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1
dev = "cuda:0"
torch.cuda.set_device(0)
sp = mdist.shard_plan(bench.VIEWS, bench.CFG, shard_of, 0)
if shard_of > 1:
    sp["emulated"] = True
pipe = bench.build_pipeline(dev)
sess, _, _ = bench.build_rank_session(dev, sp, pipe, use_graph=True)
sess.capture()
for _ in range(5):
    sess.step()
torch.cuda.synchronize()
steps = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        sess.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name and "Memcpy" not in e.name and "Memset" not in e.name]
ev.sort(key=lambda e: e.time_range.start)
if not ev:
    print("no CUDA kernel events captured (CUPTI unavailable?)")
    sys.exit(0)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy = defaultdict(float)
count = defaultdict(int)
for e in ev:
    name = e.name.split("(")[0].replace("void mvd::", "").replace("mvd::", "")
    busy[name] += e.time_range.end - e.time_range.start
    count[name] += 1
# union of busy intervals (kernels on two streams overlap)
iv = sorted((e.time_range.start, e.time_range.end) for e in ev)
covered, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > cur_e:
        covered += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
covered += cur_e - cur_s
wall = t1 - t0
print(f"shard_of={shard_of}: {len(ev) // steps} kernels/step, wall {wall / steps / 1e3:.3f} ms/step, "
      f"GPU busy (union) {covered / steps / 1e3:.3f} ms/step, idle gaps {(wall - covered) / steps / 1e3:.3f} ms/step")
print(f"{'kernel':44s} {'n/step':>7s} {'us/step':>9s} {'avg us':>8s}")
for k, v in sorted(busy.items(), key=lambda kv: -kv[1]):
    print(f"{k[:44]:44s} {count[k] / steps:7.1f} {v / steps:9.1f} {v / count[k]:8.2f}")

if "--list" in sys.argv:  # every kernel of the LAST replayed step: start (us from the step's first kernel), duration
    per = len(ev) // steps
    last = ev[-per:]
    base = last[0].time_range.start
    print(f"\nlast step, {per} kernels: start_us dur_us name")
    for e in last:
        name = e.name.split("(")[0].replace("void mvd::", "").replace("mvd::", "")
        print(f"{(e.time_range.start - base):9.1f} {(e.time_range.end - e.time_range.start):8.1f} {name[:60]}")
