"""CPU: simple-knn and Adam restatements against the reference CUDA fixture / torch.optim.Adam."""
import numpy as np
import torch

import golden_cases as gc
from oracle import raster_cpu as rc


def test_dist2_matches_reference_cuda_fixture():
    pts = gc.knn_points().numpy()
    ours = rc.dist2(pts)
    ref = gc.load("knn_5008.npz")["dist2"]
    np.testing.assert_array_equal(ours, ref)


def test_dist2_tiny_inputs():
    assert rc.dist2(np.zeros((0, 3), np.float32)).shape == (0,)
    d = rc.dist2(np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3]], np.float32))
    np.testing.assert_allclose(d, [(1 + 4 + 9) / 3, (1 + 5 + 10) / 3, (4 + 5 + 13) / 3, (9 + 10 + 13) / 3], rtol=1e-6)


def test_adam_matches_torch_cpu():
    g = torch.Generator().manual_seed(0)
    p = torch.randn(1000, 3, generator=g).requires_grad_(True)
    q = p.detach().clone().numpy()
    m = np.zeros_like(q); v = np.zeros_like(q)
    opt = torch.optim.Adam([{"params": [p], "lr": 1.6e-4}], lr=0.0, eps=1e-15)
    for step in range(1, 8):
        gr = torch.randn(1000, 3, generator=g) * (10.0 ** (step % 4 - 3))
        if step == 5:
            gr[::2] = 0
        p.grad = gr.clone()
        opt.step()
        rc.adam_step(q, gr.numpy().copy(), m, v, 1.6e-4, step)
        np.testing.assert_allclose(q, p.detach().numpy(), rtol=2e-6, atol=1e-9)
    st = opt.state[p]
    np.testing.assert_allclose(m, st["exp_avg"].numpy(), rtol=1e-6, atol=1e-30)
    np.testing.assert_allclose(v, st["exp_avg_sq"].numpy(), rtol=1e-6, atol=1e-38)
