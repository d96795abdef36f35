"""GPU: our rasterizer / distCUDA2 against the committed fixtures that the reference's own CUDA
code produced (tests/golden/*.npz) — independent of oracle/_ref being present on the box."""
import numpy as np
import pytest
import torch

import golden_cases as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["small", "medium"])
def test_rasterizer_vs_golden(name):
    from b200gs.rasterizer import _C
    c, act, cam = gc.scene(name, device="cuda")
    g = gc.load(f"raster_{name}.npz")
    bg = torch.tensor(gc.BG, device="cuda")
    E = torch.Tensor([])
    R, color, depth, radii, geom, binb, img = _C.rasterize_gaussians(
        bg, act["means3D"], E, act["opacities"], act["scales"], act["rotations"], 1.0, E, cam.viewmatrix, cam.projmatrix,
        cam.tanfovx, cam.tanfovy, c["H"], c["W"], act["shs"], 3, cam.campos, False, False)
    assert R == int(g["R"])
    np.testing.assert_array_equal(radii.cpu().numpy(), g["radii"])
    get = lambda f, dt: _C.export_state(f, c["P"], R, c["W"], c["H"], geom, binb, img).cpu().numpy().view(dt)
    np.testing.assert_array_equal(get("keys", np.uint64), g["keys"])
    np.testing.assert_array_equal(get("point_list", np.uint32), g["point_list"])
    np.testing.assert_array_equal(get("ranges", np.uint32), g["ranges"])
    np.testing.assert_array_equal(get("n_contrib", np.uint32), g["n_contrib"])
    assert np.abs(color.cpu().numpy() - g["color"]).max() <= 1e-4
    assert np.abs(depth.cpu().numpy() - g["depth"]).max() <= 1e-4
    if name == "small":
        dLc, dLd = gc.upstream(c)
        out = _C.rasterize_gaussians_backward(bg, act["means3D"], radii, E, act["scales"], act["rotations"], 1.0, E,
                                              cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, dLc.cuda(), dLd.cuda(),
                                              act["shs"], 3, cam.campos, geom, R, binb, img, False)
        names = ("means2D", "colors", "opacity", "means3D", "cov3D", "sh", "scales", "rotations")
        for k, t in zip(names, out):
            ref = g["grad_" + k].reshape(t.shape)
            rel = np.abs(t.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-30)
            assert rel < 1e-3, (k, rel)


def test_dist2_vs_golden():
    from simple_knn._C import distCUDA2
    ours = distCUDA2(gc.knn_points().cuda()).cpu().numpy()
    np.testing.assert_array_equal(ours, gc.load("knn_5008.npz")["dist2"])
