"""TEST INFRASTRUCTURE. Generates tests/golden/field_*.pt by importing the REAL reference
modules (scene/deformation.py, scene/hexplane.py from /root/reference) in this container and
running them on CPU; run here (no GPU needed):   python oracle/gen_golden_field.py
The third-party imports the reference needs but this image lacks are stubbed as SURVEY.md
Appendix E lists. Plane resolutions are kept tiny so the fixture stays small."""
import os
import sys
import types

import torch

REF = os.environ.get("REF_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    mod("tkinter", W="w")
    mod("open3d")
    mod("plyfile", PlyData=object, PlyElement=object)
    plt = mod("matplotlib.pyplot", rcParams={})
    mod("matplotlib", pyplot=plt, rcParams={})
    mod("lpips")
    mod("mmcv")
    mod("imageio")
    sk = mod("simple_knn")
    sk._C = mod("simple_knn._C", distCUDA2=lambda x: None)
    mod("diff_gaussian_rasterization", GaussianRasterizationSettings=object, GaussianRasterizer=object)


def hyper(multires, res):
    return types.SimpleNamespace(
        net_width=64, timebase_pe=4, defor_depth=0, posebase_pe=10, scale_rotation_pe=2, opacity_pe=2,
        timenet_width=64, timenet_output=32, bounds=1.6, grid_pe=0,
        kplanes_config={'grid_dimensions': 2, 'input_coordinate_dim': 4, 'output_coordinate_dim': 32, 'resolution': res},
        multires=multires, no_dx=False, no_grid=False, no_ds=False, no_dr=False, no_do=True, no_dshs=True,
        empty_voxel=False, static_mlp=False, apply_rotation=False)


def main():
    stub_modules()
    sys.path.insert(0, REF)
    from scene.deformation import deform_network           # the reference's own module
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, multires, res, P in (("l2", [1, 2], [6, 5, 7, 4], 257), ("l4", [1, 2, 4, 8], [3, 4, 3, 5], 130)):
        torch.manual_seed(6666)
        net = deform_network(hyper(multires, res))
        with torch.no_grad():
            for p in net.deformation_net.grid.grids.parameters():
                p.add_(torch.randn_like(p) * 0.05)
            for n, p in net.named_parameters():
                if n.endswith("bias"):
                    p.add_(torch.randn_like(p) * 0.05)
        net.deformation_net.set_aabb([1.4, 1.3, 1.45], [-1.35, -1.4, -1.2])
        g = torch.Generator().manual_seed(7)
        xyz = (torch.rand(P, 3, generator=g) * 3.4 - 1.7).requires_grad_(True)
        scales = (torch.randn(P, 3, generator=g) * 0.6 - 5).requires_grad_(True)
        rot = torch.randn(P, 4, generator=g).requires_grad_(True)
        opacity = torch.randn(P, 1, generator=g)
        shs = torch.randn(P, 16, 3, generator=g)
        flow = torch.randn(P, 3, generator=g) * 1e-3
        time = torch.full((P, 1), 0.37)
        frame_num = torch.tensor(22)
        pts, sc, rt, op, sh = net(xyz, scales, rot, opacity, shs, time, flow, frame_num, 1)
        feat = net.deformation_net.grid(xyz.detach(), time)
        wp, ws, wr = torch.randn(P, 3, generator=g), torch.randn(P, 3, generator=g), torch.randn(P, 4, generator=g)
        ((pts * wp).sum() + (sc * ws).sum() + (rt * wr).sum()).backward()
        grads = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
        torch.save(dict(multires=multires, res=res, state_dict={k: v.detach().clone() for k, v in net.state_dict().items()},
                        inputs=dict(xyz=xyz.detach(), scales=scales.detach(), rot=rot.detach(), opacity=opacity, shs=shs,
                                    flow=flow, time=time, frame_num=frame_num, delta_scale=1, wp=wp, ws=ws, wr=wr),
                        outputs=dict(pts=pts.detach(), scales=sc.detach(), rot=rt.detach(), feat=feat.detach()),
                        input_grads=dict(xyz=xyz.grad, scales=scales.grad, rot=rot.grad), param_grads=grads),
                   os.path.join(out_dir, f"field_{name}.pt"))
        print(name, "saved; params with grad:", len(grads))
    # regulariser: the reference's own GaussianModel.compute_regulation (scene/gaussian_model.py:730-769) on its own planes
    from scene.gaussian_model import GaussianModel
    torch.manual_seed(6666)
    net = deform_network(hyper([1, 2], [6, 5, 7, 9]))
    with torch.no_grad():
        for p in net.deformation_net.grid.grids.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    gm = GaussianModel.__new__(GaussianModel)              # only _deformation is touched by compute_regulation
    gm._deformation = net
    tw, l1w, pw = 0.01, 0.0001, 0.0001                     # arguments/dnerf/dnerf_default.py values
    loss = gm.compute_regulation(tw, l1w, pw)
    loss.backward()
    planes = {n: p.detach().clone() for n, p in net.named_parameters() if ".grids." in n}
    torch.save(dict(weights=(tw, l1w, pw), planes=planes, loss=loss.detach(),
                    grads={n: p.grad.clone() for n, p in net.named_parameters() if ".grids." in n and p.grad is not None}),
               os.path.join(out_dir, "regulation.pt"))
    print("regulation saved; loss", float(loss))


if __name__ == "__main__":
    main()
