"""Scratch timing: ours vs the reference CUDA rasterizer, fwd and bwd, CUDA events."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("iclr2025_3d-mom_b200", "iclr2025_3d-mom_b200/dropin", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import ref_harness as rh
from b200gs import synthetic as syn
from b200gs.rasterizer import _C

def ev():
    return torch.cuda.Event(enable_timing=True)

def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def run(P, W, H, mu):
    raw = syn.make_gaussians(P, scale_mu=mu, device="cuda"); act = syn.activated(raw)
    cam = syn.make_camera(W, H, device="cuda")
    bg = torch.zeros(3, device="cuda")
    E = torch.Tensor([])
    gt = torch.rand(3, H, W, device="cuda")
    out = {}
    def ours_fwd():
        out["o"] = _C.rasterize_gaussians(bg, act["means3D"], E, act["opacities"], act["scales"], act["rotations"], 1.0, E,
                                          cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, H, W, act["shs"], 3, cam.campos, False, False)
    ours_fwd()
    R, color, depth, radii, geom, binb, img = out["o"]
    dLc = torch.sign(color - gt) / (3 * H * W); dLd = torch.zeros(1, H, W, device="cuda")
    def ours_bwd():
        _C.rasterize_gaussians_backward(bg, act["means3D"], radii, E, act["scales"], act["rotations"], 1.0, E, cam.viewmatrix, cam.projmatrix,
                                        cam.tanfovx, cam.tanfovy, dLc, dLd, act["shs"], 3, cam.campos, geom, R, binb, img, False)
    L = rh.rast(); p = rh._p
    col = torch.zeros(3, H, W, device="cuda"); dep = torch.zeros(1, H, W, device="cuda"); rad = torch.zeros(P, dtype=torch.int32, device="cuda")
    def ref_fwd():
        col.zero_(); dep.zero_(); rad.zero_()
        return L.ref_rast_forward(P, 3, 16, p(bg), W, H, p(act["means3D"]), p(act["shs"]), None, p(act["opacities"]), p(act["scales"]), 1.0,
                                  p(act["rotations"]), None, p(cam.viewmatrix), p(cam.projmatrix), p(cam.campos), cam.tanfovx, cam.tanfovy, 0,
                                  p(col), p(dep), p(rad), 0)
    Rr = ref_fwd()
    def ref_bwd():
        z = lambda *s: torch.zeros(*s, device="cuda")
        g = [z(P, 3), z(P, 2, 2), z(P, 1), z(P, 3), z(P, 1), z(P, 3), z(P, 6), z(P, 16, 3), z(P, 3), z(P, 4)]
        L.ref_rast_backward(P, 3, 16, Rr, p(bg), W, H, p(act["means3D"]), p(act["shs"]), None, p(act["scales"]), 1.0, p(act["rotations"]), None,
                            p(cam.viewmatrix), p(cam.projmatrix), p(cam.campos), cam.tanfovx, cam.tanfovy, p(rad), p(dLc), p(dLd), *[p(t) for t in g], 0)
    npairs = int(torch.from_numpy(rh.ref_get("n_contrib").astype("int64")).sum())
    t = dict(ours_fwd=timeit(ours_fwd), ours_bwd=timeit(ours_bwd), ref_fwd=timeit(ref_fwd), ref_bwd=timeit(ref_bwd))
    print(f"P={P} {W}x{H} mu={mu} R={R} pairs={npairs} vis={int((radii>0).sum())} | " + " ".join(f"{k}={v:.3f}ms" for k, v in t.items())
          + f" | speedup fwd {t['ref_fwd']/t['ours_fwd']:.2f}x bwd {t['ref_bwd']/t['ours_bwd']:.2f}x total {(t['ref_fwd']+t['ref_bwd'])/(t['ours_fwd']+t['ours_bwd']):.2f}x", flush=True)

if __name__ == "__main__":
    for cfg in [(200000, 512, 512, 0.004), (200000, 512, 512, 0.010), (1000000, 1280, 720, 0.004), (1000000, 1280, 720, 0.010)]:
        run(*cfg)
