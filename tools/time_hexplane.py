"""Scratch: device time of the HexPlane kernels by plane mask (1M points, 2 levels)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("iclr2025_3d-mom_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from b200gs import field as F, engine, _lib
from b200gs._lib import check, current_stream
P = int(os.environ.get("P", 1_000_000))
torch.manual_seed(0)
net = F.deform_network(engine.default_hyper()).cuda()
grid = net.deformation_net.grid
xyz = (torch.rand(P, 3, device="cuda") * 3 - 1.5)
grid.set_aabb(xyz.max(0).values.tolist(), xyz.min(0).values.tolist())
planes = grid._planes(); levels = len(grid.grids); res = tuple(grid._res)
grads = [torch.zeros_like(p, memory_format=torch.preserve_format) for p in planes]
d = F._hex_desc(grid.aabb, planes, levels, res, grads)
order = F._cell_order(xyz, grid.aabb)
L = _lib.lib()
class _N:
    def data_ptr(self): return None
if os.environ.get("ORDER") == "identity":          # visit the (random) points in storage order: no cell coherence at all
    order = torch.arange(P, device="cuda", dtype=order.dtype)
if os.environ.get("SORTED"):
    xyz = xyz[order.long()].contiguous(); order = _N()
feat = torch.empty(P, 64, device="cuda"); S = torch.rand(P, 64, device="cuda"); A = torch.zeros(P, 64, device="cuda")
dfeat = torch.randn(P, 64, device="cuda"); dxyz = torch.empty(P, 3, device="cuda")
scratch, nb = F._time_row_scratch(d, xyz.device)
tt = torch.full((P,), 0.37, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
st = current_stream()
def fwd(mask, fac, times=None): return lambda: check(L.b200gs_hexplane_forward_masked(ctypes.byref(d), P, xyz.data_ptr(), order.data_ptr(), times.data_ptr() if times is not None else None, 0.37, mask, fac.data_ptr() if fac is not None else None, feat.data_ptr(), st))
def bwd(mask, fac, acc, rows, times=None): return lambda: check(L.b200gs_hexplane_backward_masked(ctypes.byref(d), P, xyz.data_ptr(), order.data_ptr(), times.data_ptr() if times is not None else None, 0.37, mask, fac.data_ptr() if fac is not None else None, acc.data_ptr() if acc is not None else None, dfeat.data_ptr(), dxyz.data_ptr(), scratch.data_ptr() if rows else None, nb if rows else 0, st))
print(f"fwd all           {timeit(fwd(0x3F, None)):8.1f} us")
print(f"fwd all (tensor t){timeit(fwd(0x3F, None, tt)):8.1f} us")
print(f"fwd spatial       {timeit(fwd(0x0B, None)):8.1f} us")
print(f"fwd time*S        {timeit(fwd(0x34, S)):8.1f} us")
print(f"bwd all tensor-t  {timeit(bwd(0x3F, None, None, False, tt)):8.1f} us")
print(f"bwd all scalar-t  {timeit(bwd(0x3F, None, None, False)):8.1f} us")
print(f"bwd all rows      {timeit(bwd(0x3F, None, None, True)):8.1f} us")
print(f"bwd spatial       {timeit(bwd(0x0B, None, None, False)):8.1f} us")
print(f"bwd time  classic {timeit(bwd(0x34, S, A, False)):8.1f} us")
print(f"bwd time  rows    {timeit(bwd(0x34, S, A, True)):8.1f} us")
print(f"bwd time  rows noA{timeit(bwd(0x34, S, None, True)):8.1f} us")

F_=F
def tf(): check(L.b200gs_hexplane_time_forward(ctypes.byref(d), P, xyz.data_ptr(), order.data_ptr(), 0.37, S.data_ptr(), feat.data_ptr(), 0, st))
def tb(): check(L.b200gs_hexplane_time_backward(ctypes.byref(d), P, xyz.data_ptr(), order.data_ptr(), 0.37, S.data_ptr(), A.data_ptr(), dfeat.data_ptr(), dxyz.data_ptr(), scratch.data_ptr(), nb, 0, st))
print(f"time-rows fwd     {timeit(tf):8.1f} us")
print(f"time-rows bwd     {timeit(tb):8.1f} us")
# agreement with the masked kernels
f1 = torch.empty_like(feat); f2 = torch.empty_like(feat)
check(L.b200gs_hexplane_forward_masked(ctypes.byref(d), P, xyz.data_ptr(), order.data_ptr(), None, 0.37, 0x34, S.data_ptr(), f1.data_ptr(), st))
check(L.b200gs_hexplane_time_forward(ctypes.byref(d), P, xyz.data_ptr(), order.data_ptr(), 0.37, S.data_ptr(), f2.data_ptr(), 0, st))
print("fwd max rel diff", ((f1 - f2).abs().max() / f1.abs().max()).item())
