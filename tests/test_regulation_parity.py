"""GPU parity of the fused HexPlane regulariser against the oracle restatement of the reference's formulas
(oracle/field_torch.py::plane_regulation = scene/gaussian_model.py:730-769 + scene/regulation.py:22-28, pinned bit-exact
against the real GaussianModel.compute_regulation by tests/golden/regulation.pt) and against that golden fixture itself.
Value within 1e-5 relative, plane gradients within 1e-5 of each plane's largest gradient."""
import os

import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(grids, tw, l1w, pw):
    from oracle import field_torch
    return field_torch.plane_regulation(grids, tw, l1w, pw)


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


def test_regulation_against_reference_golden():
    """The kernel on the very planes the reference's own GaussianModel.compute_regulation was run on."""
    from b200gs import field
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regulation.pt"))
    cfg = {'grid_dimensions': 2, 'input_coordinate_dim': 4, 'output_coordinate_dim': 32, 'resolution': [6, 5, 7, 9]}
    f = field.HexPlaneField(1.6, cfg, [1, 2]).cuda()
    with torch.no_grad():
        for l, gp in enumerate(f.grids):
            for k, p in enumerate(gp):
                p.copy_(g["planes"][f"deformation_net.grid.grids.{l}.{k}"].cuda())
    loss = field.compute_regulation(f, *g["weights"])
    loss.backward()
    assert abs(loss.item() - g["loss"].item()) < 1e-5 * abs(g["loss"].item())
    for l, gp in enumerate(f.grids):
        for k, p in enumerate(gp):
            ref = g["grads"][f"deformation_net.grid.grids.{l}.{k}"].cuda()
            assert (p.grad - ref).abs().max().item() <= 1e-5 * ref.abs().max().item(), (l, k)
