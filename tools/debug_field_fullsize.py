import os, sys
sys.path.insert(0, "/root/repo/iclr2025_3d-mom_b200"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import torch
import test_field_parity as T
from oracle import field_torch as oracle
P = int(os.environ.get("P", 1000000))
net = T._model([1, 2], 50)
xyz, scales, rot, opacity, shs, flow = T._inputs(P)
time = torch.full((P, 1), 0.37, device="cuda"); frame_num = torch.tensor(22, device="cuda")
sd = {k: v.detach().clone().contiguous().requires_grad_(v.dtype.is_floating_point) for k, v in net.state_dict().items()}
a = [t.clone().requires_grad_(True) for t in (xyz, scales, rot)]
pts, sc, rt, op, sh = net(a[0], a[1], a[2], opacity, shs, time, flow, frame_num, 1)
b = [t.clone().requires_grad_(True) for t in (xyz, scales, rot)]
rp, rs, rr, ro, rsh = oracle.deform_forward(sd, 2, b[0], b[1], b[2], opacity, shs, time, flow, frame_num, 1)
g = torch.Generator().manual_seed(5)
wp, ws, wr = (torch.randn(P, 3, generator=g).cuda(), torch.randn(P, 3, generator=g).cuda(), torch.randn(P, 4, generator=g).cuda())
((pts * wp).sum() + (sc * ws).sum() + (rt * wr).sum()).backward()
((rp * wp).sum() + (rs * ws).sum() + (rr * wr).sum()).backward()
d = (a[0].grad - b[0].grad).abs()
mx = b[0].grad.abs().max().item()
print("max ref grad", mx, "max diff", d.max().item(), "n > 1e-3*max:", int((d > 1e-3 * mx).sum()), "n > 1e-4*max:", int((d > 1e-4 * mx).sum()))
idx = torch.nonzero(d.max(dim=1).values > 1e-3 * mx).flatten()[:8]
aabb = net.deformation_net.grid.aabb
for i in idx.tolist():
    n = (xyz[i] - aabb[0]) * (2.0 / (aabb[1] - aabb[0])) - 1.0
    print(i, "xyz", xyz[i].tolist(), "ours", a[0].grad[i].tolist(), "ref", b[0].grad[i].tolist())
    for res in (64, 128):
        ix = ((n + 1) / 2) * (res - 1)
        print("    res", res, "pixel coords", ix.tolist(), "frac", (ix - ix.floor()).tolist())
