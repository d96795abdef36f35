"""Scratch: torch.profiler (CUPTI) kernel table for one bench-style training step (warm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from torch.profiler import profile, ProfilerActivity

class A: pass
args = A(); args.points = int(os.environ.get("P", 1_000_000)); args.width = 1280; args.height = 720
args.views_per_gpu = int(os.environ.get("V", 4)); args.scale_mu = 0.01
dev = torch.device("cuda", 0)
raw, cams, gts_host, n_global = bench.build_scene(args, dev, 1, 0, "b200")
model, trainer = bench.make_b200_trainer(args, raw, dev, 1, 0)
gts = [g.to(dev) for g in gts_host]
for _ in range(3):
    trainer.step(cams, gts, global_batch=n_global)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        trainer.step(cams, gts, global_batch=n_global)
    torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev) / 1e3
agg = {}
for e in ev:
    k = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
    a = agg.setdefault(k, [0.0, 0]); a[0] += e.device_time / 1e3; a[1] += 1
nviews = 2 * args.views_per_gpu
print(f"wall {wall:.1f} ms for {nviews} views (under profiler); GPU kernel time {tot:.1f} ms; per view {tot/nviews:.3f} ms")
for k, (v, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:28]:
    print(f"{v/nviews*1e3:9.1f} us/view {100*v/tot:5.1f}%  x{n/nviews:5.1f}  {k}")
