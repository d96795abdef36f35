"""Scratch: weight-gradient error of the fused field vs torch FP32 and FP64 oracles at 1M points."""
import os, sys
sys.path.insert(0, "/root/repo/iclr2025_3d-mom_b200"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import torch
import test_field_parity as T
from oracle import field_torch as oracle
P = int(os.environ.get("P", 1000000))
net = T._model([1, 2], 50)
xyz, scales, rot, opacity, shs, flow = T._inputs(P)
time = torch.full((P, 1), 0.37, device="cuda"); frame_num = torch.tensor(22, device="cuda")
g = torch.Generator().manual_seed(5)
wp, ws, wr = (torch.randn(P, 3, generator=g).cuda(), torch.randn(P, 3, generator=g).cuda(), torch.randn(P, 4, generator=g).cuda())
def run_oracle(dtype):
    sd = {k: (v.detach().clone().to(dtype) if v.dtype.is_floating_point else v.detach().clone()).contiguous().requires_grad_(v.dtype.is_floating_point) for k, v in net.state_dict().items()}
    b = [t.clone().to(dtype).requires_grad_(True) for t in (xyz, scales, rot)]
    rp, rs, rr, ro, rsh = oracle.deform_forward(sd, 2, b[0], b[1], b[2], opacity.to(dtype), shs.to(dtype), time.to(dtype), flow.to(dtype), frame_num, 1)
    ((rp * wp.to(dtype)).sum() + (rs * ws.to(dtype)).sum() + (rr * wr.to(dtype)).sum()).backward()
    return sd
a = [t.clone().requires_grad_(True) for t in (xyz, scales, rot)]
pts, sc, rt, op, sh = net(a[0], a[1], a[2], opacity, shs, time, flow, frame_num, 1)
((pts * wp).sum() + (sc * ws).sum() + (rt * wr).sum()).backward()
s32, s64 = run_oracle(torch.float32), run_oracle(torch.float64)
params = dict(net.named_parameters())
for k in ["deformation_net.rotations_deform.1.weight", "deformation_net.pos_deform.1.weight", "deformation_net.feature_out.0.weight",
          "deformation_net.rotations_deform.3.weight", "deformation_net.grid.grids.0.1", "deformation_net.grid.grids.1.0"]:
    o, r32, r64 = params[k].grad.double(), s32[k].grad.double(), s64[k].grad
    if o.dim() == 4: o = o.contiguous()
    mx = r64.abs().max()
    print(f"{k:48s} |max| {mx.item():.3e}  ours-fp64 {((o - r64).abs().max() / mx).item():.2e}  torch32-fp64 {((r32 - r64).abs().max() / mx).item():.2e}  ours-torch32 {((o - r32).abs().max() / mx).item():.2e}")
