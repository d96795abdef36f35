"""GPU: the trainer's fast path (shared per-step SH tensor with its gradient accumulated by the rasterizer backward into one
[P,16,3] buffer, fused L1, field kernels adding straight into the flat arena, SH Adam on the side stream from that buffer) against
the SAME kernels driven through plain autograd (per-view torch.cat of the SH tensors, autograd accumulation, (render - gt).abs().mean(),
one FusedAdam.step over everything) -- for the coarse stage (no deformation field, train_4DGS.py's first stage) and the fine stage.
Same model, same views: losses equal, every gradient equal to float-atomic order, parameters after two steps equal except where
|g| is rounding noise."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("stage,views", [("coarse", 2), ("fine", 3), ("fine", 1)])
def test_fast_path_equals_plain_autograd_path(stage, views):
    from b200gs import engine, synthetic as syn
    dev = torch.device("cuda", 0)
    P, W, H = 30000, 240, 144
    raw = syn.make_gaussians(P, scale_mu=0.012, seed=3, device="cpu")
    torch.manual_seed(0)
    ma = engine.GaussianState({k: v.to(dev) for k, v in raw.items()}).to(dev)
    with torch.no_grad():
        for p in ma._deformation.deformation_net.grid.grids.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    mb = copy.deepcopy(ma)
    ma.training_setup(); mb.training_setup()
    bg = torch.tensor([0.05, 0.1, 0.15], device=dev)
    h = engine.default_hyper()
    reg = (h.time_smoothness_weight, h.l1_time_planes, h.plane_tv_weight)
    ta = engine.ViewParallelTrainer(ma, bg, stage=stage, regulation=reg)
    tb = engine.ViewParallelTrainer(mb, bg, stage=stage, regulation=reg, shared_shs=False,
                                    render_fn=lambda cam, m, b, st: engine.render(cam, m, b, stage=st))
    assert len(ta.sh_params) == 2 and len(tb.sh_params) == 0
    cams = syn.orbit_cameras(views, W, H, device=dev)
    g = torch.Generator().manual_seed(1)
    gts = [torch.rand(3, H, W, generator=g).to(dev) for _ in cams]
    for it in range(2):
        la = float(ta.step(cams, gts)); lb = float(tb.step(cams, gts))
        assert abs(la - lb) <= 2e-6 * max(1.0, abs(lb)), (it, la, lb)
        # utils/image_utils.py:psnr as train_4DGS.py:212 logs it: the L1 kernel's side sum against the plain torch expression
        pa, pb = float(ta.psnr()), float(tb.psnr())
        assert abs(pa - pb) <= 1e-4 * abs(pb), (pa, pb)
        if it == 0:
            na = {n: p for n, p in ma.named_parameters()}
            for n, q in mb.named_parameters():
                p = na[n]
                if q.grad is None:
                    assert p.grad is None or n in ("_features_dc", "_features_rest"), n
                    continue
                gp = p.grad
                if n == "_features_dc":
                    gp = ta.sh_grad[:, :1]
                elif n == "_features_rest":
                    gp = ta.sh_grad[:, 1:]
                assert gp is not None, n
                if float(q.grad.abs().max()) == 0.0:
                    assert float(gp.abs().max()) == 0.0, n
                else:
                    assert _rel(gp, q.grad) <= 2e-5, (n, _rel(gp, q.grad))
            assert _rel(ta.viewspace_grad, tb.viewspace_grad) <= 2e-5
            assert torch.equal(ta.max_radii, tb.max_radii)
    if stage == "coarse":      # the deformation field is bypassed: its parameters never receive a gradient and never move
        assert all(p.grad is None for n, p in ma.named_parameters() if n.startswith("_deformation."))
    for (n, p), (_, q) in zip(ma.named_parameters(), mb.named_parameters()):
        d = (p.detach() - q.detach()).abs()
        assert (d > 1e-6).float().mean().item() < 2e-3, (n, d.max().item())
