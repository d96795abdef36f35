import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
os.environ["V"] = "4"
import torch
from torch.profiler import profile, ProfilerActivity
import bench
class A: pass
args = A(); args.points = 1_000_000; args.width = 1280; args.height = 720; args.views_per_gpu = 4; args.scale_mu = 0.01
dev = torch.device("cuda", 0)
raw, cams, gts_host, n_global = bench.build_scene(args, dev, 1, 0, "b200")
model, trainer = bench.make_b200_trainer(args, raw, dev, 1, 0)
gts = [g.to(dev) for g in gts_host]
for _ in range(3): trainer.step(cams, gts, global_batch=n_global)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    trainer.step(cams, gts, global_batch=n_global); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.device_time_total > 0 and not e.key.startswith("b200gs") and "Memcpy" not in e.key:
        rows.append((e.device_time_total, e.count, e.key[:60], str(e.input_shapes)[:90]))
for t, c, k, s in sorted(rows, reverse=True)[:22]:
    print(f"{t/4:8.1f} us/view x{c/4:5.1f} {k} {s}")
