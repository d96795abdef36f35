"""Scratch: standalone HexPlaneField vs the torch oracle at 1M points: where do per-point xyz gradients differ?"""
import os, sys
sys.path.insert(0, "/root/repo/iclr2025_3d-mom_b200"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import torch
from oracle import field_torch as oracle
from b200gs.field import HexPlaneField
torch.manual_seed(3)
cfg = {'grid_dimensions': 2, 'input_coordinate_dim': 4, 'output_coordinate_dim': 32, 'resolution': [64, 64, 64, 50]}
f = HexPlaneField(1.6, cfg, [1, 2]).cuda()
with torch.no_grad():
    for p in f._planes(): p.add_(torch.randn_like(p) * 0.01)
f.set_aabb([1.4, 1.3, 1.45], [-1.35, -1.4, -1.2])
P = int(os.environ.get("P", 1000000))
g = torch.Generator().manual_seed(4)
pts = (torch.rand(P, 3, generator=g) * 3.4 - 1.7).cuda().requires_grad_(True)
t = torch.full((P, 1), 0.37).cuda()
feat = f(pts, t)
planes = [p.detach().clone().contiguous().requires_grad_(True) for gp in f.grids for p in gp]
pts2 = pts.detach().clone().requires_grad_(True)
ref = oracle.hexplane_features(pts2, t, f.aabb.detach(), planes, 2)
print("fwd rel", ((feat - ref).abs().max() / ref.abs().max()).item())
w = torch.randn(P, 64, generator=g).cuda()
(feat * w).sum().backward(); (ref * w).sum().backward()
d = (pts.grad - pts2.grad).abs(); mx = pts2.grad.abs().max().item()
print("xyz grad: max ref", mx, "max diff", d.max().item(), "n>1e-3", int((d > 1e-3 * mx).sum()), "n>1e-4", int((d > 1e-4 * mx).sum()))
for pa, pb in zip(f._planes(), planes):
    pass
idx = torch.nonzero(d.max(dim=1).values > 1e-3 * mx).flatten()[:6]
for i in idx.tolist():
    print(i, pts[i].tolist(), "ours", pts.grad[i].tolist(), "ref", pts2.grad[i].tolist())
# is torch deterministic / self-consistent?  recompute the oracle gradient in float64
pts3 = pts.detach().double().requires_grad_(True)
planes64 = [p.detach().double() for p in planes]
ref64 = oracle.hexplane_features(pts3, t.double(), f.aabb.detach().double(), planes64, 2)
(ref64 * w.double()).sum().backward()
d_o = (pts.grad.double() - pts3.grad).abs(); d_r = (pts2.grad.double() - pts3.grad).abs()
print("vs float64 oracle: ours max diff", d_o.max().item(), " torch-fp32 max diff", d_r.max().item(), " n(ours>1e-3)", int((d_o > 1e-3 * mx).sum()), " n(torch32>1e-3)", int((d_r > 1e-3 * mx).sum()))
