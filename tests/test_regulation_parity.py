"""GPU parity of the fused HexPlane regulariser against the reference's formulas
(scene/gaussian_model.py:730-769 _plane_regulation / _time_regulation / _l1_regulation,
scene/regulation.py:22-28 compute_plane_smoothness), restated here with the same PyTorch ops.
Value within 1e-5 relative, plane gradients within 1e-5 of each plane's largest gradient."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _smooth(t):
    h = t.shape[2]
    first = t[..., 1:, :] - t[..., :h - 1, :]
    second = first[..., 1:, :] - first[..., :h - 2, :]
    return torch.square(second).mean()


def _reference(grids, tw, l1w, pw):
    plane = sum(_smooth(g[k]) for g in grids for k in (0, 1, 3))
    time = sum(_smooth(g[k]) for g in grids for k in (2, 4, 5))
    l1 = sum(torch.abs(1 - g[k]).mean() for g in grids for k in (2, 4, 5))
    return pw * plane + tw * time + l1w * l1


@pytest.mark.parametrize("multires,T", [([1, 2], 50), ([1, 2, 4, 8], 25)])
def test_regulation_value_and_gradient(multires, T):
    from b200gs import field
    torch.manual_seed(7)
    cfg = {'grid_dimensions': 2, 'input_coordinate_dim': 4, 'output_coordinate_dim': 32, 'resolution': [64, 64, 64, T]}
    f = field.HexPlaneField(1.6, cfg, multires).cuda()
    with torch.no_grad():
        for p in f._planes():
            p.add_(torch.randn_like(p) * 0.05)
    tw, l1w, pw = 0.01, 0.0001, 0.0001
    ref_grids = [[p.detach().clone().contiguous().requires_grad_(True) for p in gp] for gp in f.grids]
    ref = _reference(ref_grids, tw, l1w, pw)
    ref.backward()
    ours = field.compute_regulation(f, tw, l1w, pw)
    (ours * 3.0).backward()
    assert abs(ours.item() - ref.item()) < 1e-5 * abs(ref.item()), (ours.item(), ref.item())
    for gp, rg in zip(f.grids, ref_grids):
        for p, q in zip(gp, rg):
            assert (p.grad / 3.0 - q.grad).abs().max().item() <= 1e-5 * q.grad.abs().max().item()
    # accumulate path: adds on top of existing .grad buffers and into the loss accumulator
    before = [p.grad.clone() for p in f._planes()]
    acc = torch.zeros(1, device="cuda")
    field.accumulate_regulation(f, tw, l1w, pw, loss_accum=acc)
    assert abs(acc.item() - ref.item()) < 1e-5 * abs(ref.item())
    for p, b, q in zip(f._planes(), before, [q for rg in ref_grids for q in rg]):
        assert ((p.grad - b) - q.grad).abs().max().item() <= 1e-5 * q.grad.abs().max().item()
