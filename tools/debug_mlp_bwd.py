"""Scratch: per-parameter gradient error of the fused field vs the torch oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("iclr2025_3d-mom_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import test_field_parity as T
from oracle import field_torch as oracle
P = int(os.environ.get("P", 5000))
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
torch.cuda.synchronize()
for x, y, n in zip(a, b, ("xyz", "scales", "rot")):
    print(f"{n:50s} {T._rel(x.grad, y.grad):.3e}")
params = dict(net.named_parameters())
torch.set_printoptions(precision=4, linewidth=200)
for k, v in sd.items():
    if k in params and v.grad is not None and params[k].grad is not None and "grids" not in k:
        pg = params[k].grad
        d = (pg - v.grad).flatten(); r = v.grad.flatten()
        big = r.abs() > 0.3 * r.abs().max()
        print(f"{k:50s} {T._rel(pg, v.grad):.3e}  ref|max|={r.abs().max().item():.3e}  rms_err/rms={d.pow(2).mean().sqrt().item()/r.pow(2).mean().sqrt().item():.2e}  mean signed rel err on big elems={(d[big]/r[big]).mean().item():+.2e} (n={int(big.sum())})")
        if T._rel(pg, v.grad) > 1e-3 and os.environ.get("SHOW"):
            print(" ours", pg.flatten()[:8].tolist()); print(" ref ", v.grad.flatten()[:8].tolist())
