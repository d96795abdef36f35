"""GPU: the plane-subset HexPlane kernels against the six-plane kernels they re-associate.

features = S * T with S = spatial planes (mask 0x0B, evaluated once per step) and T = time planes (mask 0x34, per view)
must equal the six-plane product up to FP32 re-association (2e-6 relative), and the split backward (time pass with the
gradient accumulator + one deferred spatial pass) must reproduce the six-plane backward's plane and xyz gradients (1e-4
of each tensor's largest gradient; the accumulations are re-ordered). Covers the three time-plane gradient routes: classic
four-texel REDs (per-point times), replicated 1-D rows (scalar time) and the shared-memory row kernels."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("P", [1, 4099, 50001])
def test_split_equals_six_plane(P):
    from b200gs import field as F, engine, _lib
    from b200gs._lib import check, current_stream
    torch.manual_seed(5)
    net = F.deform_network(engine.default_hyper()).cuda()
    grid = net.deformation_net.grid
    with torch.no_grad():
        for p in grid._planes():
            p.add_(torch.randn_like(p) * 0.05)
    xyz = (torch.rand(P, 3, device="cuda") * 3.6 - 1.8)          # some points outside the box -> border clamp
    grid.set_aabb([1.5, 1.4, 1.45], [-1.5, -1.45, -1.4])
    planes, levels, res = grid._planes(), len(grid.grids), tuple(grid._res)
    L, st, t = _lib.lib(), current_stream(), 0.37
    order = F._cell_order(xyz, grid.aabb)
    optr = order.data_ptr() if order is not None else None
    new_grads = lambda: [torch.zeros_like(p, memory_format=torch.preserve_format) for p in planes]
    dfeat = torch.randn(P, 64, device="cuda")
    dfeat[::3] = 0                                                 # rows without gradient are skipped by the kernels

    # six planes, per-point times (classic route)
    g_full = new_grads(); d_full = F._hex_desc(grid.aabb, planes, levels, res, g_full)
    tt = torch.full((P,), t, device="cuda")
    f_full = torch.empty(P, 64, device="cuda"); dx_full = torch.empty(P, 3, device="cuda")
    check(L.b200gs_hexplane_forward(ctypes.byref(d_full), P, xyz.data_ptr(), optr, tt.data_ptr(), 0.0, f_full.data_ptr(), st))
    check(L.b200gs_hexplane_backward(ctypes.byref(d_full), P, xyz.data_ptr(), optr, tt.data_ptr(), 0.0, dfeat.data_ptr(), dx_full.data_ptr(), st))

    def split(route):
        g = new_grads(); d = F._hex_desc(grid.aabb, planes, levels, res, g)
        S = torch.empty(P, 64, device="cuda"); A = torch.zeros(P, 64, device="cuda")
        f = torch.empty(P, 64, device="cuda"); dx_t = torch.empty(P, 3, device="cuda"); dx_s = torch.empty(P, 3, device="cuda")
        check(L.b200gs_hexplane_forward_masked(ctypes.byref(d), P, xyz.data_ptr(), optr, None, 0.0, F.MASK_SPATIAL, None, S.data_ptr(), st))
        scratch, nb = F._time_row_scratch(d, xyz.device)
        if route == "smem":
            assert L.b200gs_hexplane_time_supported(ctypes.byref(d)) == 1
            check(L.b200gs_hexplane_time_forward(ctypes.byref(d), P, xyz.data_ptr(), None, t, S.data_ptr(), f.data_ptr(), 0, st))
            check(L.b200gs_hexplane_time_backward(ctypes.byref(d), P, xyz.data_ptr(), None, t, S.data_ptr(), A.data_ptr(), dfeat.data_ptr(),
                                                  dx_t.data_ptr(), scratch.data_ptr(), nb, 0, st))
        else:
            times = tt.data_ptr() if route == "classic" else None
            check(L.b200gs_hexplane_forward_masked(ctypes.byref(d), P, xyz.data_ptr(), optr, times, t, F.MASK_TIME, S.data_ptr(), f.data_ptr(), st))
            check(L.b200gs_hexplane_backward_masked(ctypes.byref(d), P, xyz.data_ptr(), optr, times, t, F.MASK_TIME, S.data_ptr(), A.data_ptr(),
                                                    dfeat.data_ptr(), dx_t.data_ptr(), scratch.data_ptr() if route == "rows" else None,
                                                    nb if route == "rows" else 0, st))
        check(L.b200gs_hexplane_backward_masked(ctypes.byref(d), P, xyz.data_ptr(), optr, None, 0.0, F.MASK_SPATIAL, None, None, A.data_ptr(),
                                                dx_s.data_ptr(), None, 0, st))
        return f, dx_t + dx_s, g

    for route in ("classic", "rows", "smem"):
        f, dx, g = split(route)
        assert _rel(f, f_full) < 2e-6, route
        assert _rel(dx, dx_full) < 1e-4, route
        for a, b in zip(g, g_full):
            assert _rel(a, b) < 1e-4, route
