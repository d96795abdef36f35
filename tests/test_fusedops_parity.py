"""GPU parity of the fused elementwise kernels against the PyTorch ops the reference uses:
exp / F.normalize / sigmoid (gaussian_renderer/__init__.py:130-132) forward + backward, and
l1_loss (utils/loss_utils.py:23-24) with its gradient. Tolerance: 2e-6 relative (same FP32 formulas,
different instruction order); the loss sum is accumulated in a different order (1e-5)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("P", [1, 257, 100003])
def test_activations_forward_backward(P):
    from b200gs import fusedops
    g = torch.Generator().manual_seed(P)
    s = (torch.randn(P, 3, generator=g) * 0.6 - 5).cuda()
    r = torch.randn(P, 4, generator=g).cuda()
    o = (torch.randn(P, 1, generator=g) * 1.5).cuda()
    if P > 2:
        r[1] = 0.0                       # zero quaternion: the eps clamp of F.normalize is active
    a = [t.clone().requires_grad_(True) for t in (s, r, o)]
    b = [t.clone().requires_grad_(True) for t in (s, r, o)]
    so, ro, oo = fusedops.activations(*a)
    rs, rr, ropa = torch.exp(b[0]), torch.nn.functional.normalize(b[1]), torch.sigmoid(b[2])
    assert _rel(so, rs) < 2e-6 and _rel(ro, rr) < 2e-6 and _rel(oo, ropa) < 2e-6
    ws, wr, wo = torch.randn(P, 3, generator=g).cuda(), torch.randn(P, 4, generator=g).cuda(), torch.randn(P, 1, generator=g).cuda()
    ((so * ws).sum() + (ro * wr).sum() + (oo * wo).sum()).backward()
    ((rs * ws).sum() + (rr * wr).sum() + (ropa * wo).sum()).backward()
    live = torch.ones(P, dtype=torch.bool, device="cuda")
    if P > 2:
        live[1] = False                  # d normalize at exactly 0 is 1e12 * g in both; compare separately
        assert torch.allclose(a[1].grad[1], b[1].grad[1], rtol=1e-5)
    for x, y in zip(a, b):
        assert _rel(x.grad[live], y.grad[live]) < 2e-6


@pytest.mark.parametrize("shape", [(3, 720, 1280), (3, 5, 7)])
def test_l1_loss_and_grad(shape):
    from b200gs import fusedops
    g = torch.Generator().manual_seed(3)
    img = torch.rand(*shape, generator=g).cuda().requires_grad_(True)
    gt = torch.rand(*shape, generator=g).cuda()
    with torch.no_grad():
        gt.view(-1)[:3] = img.view(-1)[:3]          # exact ties: sign(0) = 0
    B = 8
    acc, sse = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    d = fusedops.l1_loss_and_grad(img, gt, 1.0 / (img.numel() * B), acc, sse)
    ref = (img - gt).abs().mean() / B
    ref.backward()
    assert abs(acc.item() - ref.item()) < 1e-5 * ref.item()
    assert torch.equal(d, img.grad)
    # utils/image_utils.py:17-38 (no mask): mse = ((a - b) ** 2).view(B, -1).mean(1); psnr = 20 log10(1 / sqrt(mse))
    a, b = img.detach()[None], gt[None]
    mse = ((a - b) ** 2).view(a.shape[0], -1).mean(1, keepdim=True)
    want = 20 * torch.log10(1.0 / torch.sqrt(mse))
    got = fusedops.psnr_from_sse(sse, img.numel())
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
    assert torch.equal(fusedops.l1_loss_and_grad(img, gt, 1.0 / (img.numel() * B), acc), d)      # sse stays optional


def test_to8b_and_frame_ring():
    import numpy as np
    from b200gs import output
    g = torch.Generator().manual_seed(11)
    frames = [(torch.rand(3, 67, 129, generator=g) * 1.4 - 0.2).cuda() for _ in range(6)]
    to8b = lambda x: (255 * np.clip(x.cpu().numpy(), 0, 1)).astype(np.uint8)          # render_4DGS.py:49
    for f in frames[:2]:
        assert np.array_equal(output.to8b(f).cpu().numpy(), to8b(f).transpose(1, 2, 0))
    ring = output.FrameRing(67, 129, depth=3)
    got = []
    for i, f in enumerate(frames):
        if ring.count == 3:
            got.append(ring.pop().copy())
        ring.push(f)
    while ring.count:
        got.append(ring.pop().copy())
    assert len(got) == 6 and all(np.array_equal(a, to8b(f).transpose(1, 2, 0)) for a, f in zip(got, frames))
    # the whole path in one pinned array, frames written out of order, refilled for a second path
    store = output.FrameStore(6, 67, 129)
    for i in (3, 0, 5, 1, 4, 2):
        store.put(i, frames[i])
    arr = store.array()
    assert arr.shape == (6, 67, 129, 3) and all(np.array_equal(arr[i], to8b(f).transpose(1, 2, 0)) for i, f in enumerate(frames))
    for i in range(6):
        store.put(i, frames[5 - i])
    arr = store.array()
    assert all(np.array_equal(arr[i], to8b(frames[5 - i]).transpose(1, 2, 0)) for i in range(6))
    with pytest.raises(IndexError):
        store.put(6, frames[0])


def test_l1_against_uint8_target_equals_the_float_path():
    """The dataset's own uint8 HWC image as the L1 target (converted on the device as utils/general_utils.py:PILtoTorch converts
    it on the host: u8 / 255.0) gives the same gradient bits and the same loss as the float CHW target."""
    from b200gs import fusedops
    g = torch.Generator().manual_seed(5)
    H, W = 75, 130
    u8 = (torch.rand(H, W, 3, generator=g) * 255.999).to(torch.uint8)
    target = (u8.float() / 255.0).permute(2, 0, 1).contiguous().cuda()
    render = torch.rand(3, H, W, generator=g).cuda()
    render[0, :4, :4] = target[0, :4, :4]                     # exact hits: sign(0) = 0
    la, lb = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    sa, sb = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    da = fusedops.l1_loss_and_grad(render, target, 0.125 / render.numel(), la, sa)
    db = fusedops.l1_loss_and_grad(render, u8.cuda(), 0.125 / render.numel(), lb, sb)
    assert torch.equal(da, db)
    assert abs(float(sa) - float(sb)) <= 1e-6 * float(sa)
    assert abs(float(sb) - float(((render - target) ** 2).sum())) <= 1e-5 * float(sb)
    assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))
    ref = 0.125 * (render - target).abs().mean()
    assert abs(float(lb) - float(ref)) <= 2e-6 * float(ref)


def test_host_image_feeder_and_loss_ring():
    """engine.HostImageFeeder (ground truth uploaded one step ahead on a side stream) hands every step the right images;
    engine.LossRing returns the value pushed `lag` steps ago without a per-step device sync."""
    from b200gs import engine
    dev = torch.device("cuda", 0)
    steps = [[(torch.full((8, 9, 3), 10 * s + v, dtype=torch.uint8)).pin_memory() for v in range(2)] for s in range(5)]
    feeder = engine.HostImageFeeder(steps[0], dev)
    ring = engine.LossRing(depth=4)
    feeder.prefetch(steps[0])
    for s in range(5):
        imgs = feeder.take()
        if s + 1 < 5:
            feeder.prefetch(steps[s + 1])
        got = [int(t.float().mean().item()) for t in imgs]
        assert got == [10 * s, 10 * s + 1], (s, got)
        torch.cuda._sleep(200000)                       # the step's work
        feeder.release()
        ring.push(torch.tensor([float(s)], device=dev))
        assert ring.read(lag=1) == (None if s == 0 else float(s - 1))
    assert ring.read(lag=0) == 4.0
