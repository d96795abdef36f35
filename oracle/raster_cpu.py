"""TEST INFRASTRUCTURE — NOT PRODUCT CODE. numpy/ctypes front-end of oracle/raster_cpu.c."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def build():
    src = os.path.join(HERE, "raster_cpu.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_bin.restype = ctypes.c_longlong
    return _lib


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def forward(means, opac, view, proj, campos, W, H, tanx, tany, bg, shs=None, colors_pre=None, scales=None, rots=None,
            cov3d_pre=None, D=3, mod=1.0, composite=True):
    """Reference forward on numpy arrays. view / proj are the TRANSPOSED 4x4 matrices the
    rasterizer receives (flattened row-major = column-major of the maths matrix)."""
    L = lib()
    means = _f(means); opac = _f(opac).reshape(-1); shs = _f(shs); colors_pre = _f(colors_pre)
    scales = _f(scales); rots = _f(rots); cov3d_pre = _f(cov3d_pre)
    view = _f(view).reshape(-1); proj = _f(proj).reshape(-1); campos = _f(campos); bg = _f(bg)
    P = means.shape[0]
    M = shs.shape[1] if shs is not None else 0
    s = dict(radii=np.zeros(P, np.int32), xy=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
             cov3D=np.zeros((P, 6), np.float32), rgb=np.zeros((P, 3), np.float32),
             conic_opacity=np.zeros((P, 4), np.float32), tiles_touched=np.zeros(P, np.uint32),
             clamped=np.zeros((P, 3), np.uint8), rect=np.zeros((P, 4), np.uint32))
    L.orc_preprocess(P, D, M, _p(means), _p(scales), ctypes.c_float(mod), _p(rots), _p(opac), _p(shs), _p(cov3d_pre),
                     _p(colors_pre), _p(view), _p(proj), _p(campos), W, H, ctypes.c_float(tanx), ctypes.c_float(tany),
                     _p(s["radii"]), _p(s["xy"]), _p(s["depths"]), _p(s["cov3D"]), _p(s["rgb"]), _p(s["conic_opacity"]),
                     _p(s["tiles_touched"]), _p(s["clamped"]), _p(s["rect"]))
    R = int(s["tiles_touched"].astype(np.int64).sum())
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    s.update(R=R, keys=np.zeros(R, np.uint64), point_list=np.zeros(R, np.uint32), ranges=np.zeros((tiles, 2), np.uint32))
    got = L.orc_bin(P, W, _p(s["radii"]), _p(s["depths"]), _p(s["tiles_touched"]), _p(s["rect"]), ctypes.c_longlong(R),
                    _p(s["keys"]), _p(s["point_list"]), _p(s["ranges"]), tiles)
    assert got == R
    if composite:
        s.update(final_T=np.zeros(H * W, np.float32), n_contrib=np.zeros(H * W, np.uint32),
                 color=np.zeros((3, H, W), np.float32), depth=np.zeros((1, H, W), np.float32))
        feat = colors_pre if colors_pre is not None else s["rgb"]
        L.orc_render_fwd(W, H, _p(s["ranges"]), _p(s["point_list"]), _p(s["xy"]), _p(feat), _p(s["depths"]),
                         _p(s["conic_opacity"]), _p(bg), _p(s["final_T"]), _p(s["n_contrib"]), _p(s["color"]), _p(s["depth"]))
    s["inputs"] = dict(means=means, opac=opac, shs=shs, colors_pre=colors_pre, scales=scales, rots=rots,
                       cov3d_pre=cov3d_pre, view=view, proj=proj, campos=campos, bg=bg, W=W, H=H, tanx=tanx, tany=tany,
                       D=D, M=M, mod=mod)
    return s


def backward(s, dL_dpix, dL_ddepth=None):
    """Reference backward from the state of `forward`. Returns the gradient dict."""
    L = lib()
    i = s["inputs"]
    P = i["means"].shape[0]; W, H, M = i["W"], i["H"], i["M"]
    dL_dpix = _f(dL_dpix); dL_ddepth = _f(dL_ddepth)
    g = dict(means2D=np.zeros((P, 3), np.float32), conic=np.zeros((P, 4), np.float32), opacity=np.zeros(P, np.float32),
             colors=np.zeros((P, 3), np.float32), depths=np.zeros(P, np.float32), means3D=np.zeros((P, 3), np.float32),
             cov3D=np.zeros((P, 6), np.float32), sh=np.zeros((P, M, 3), np.float32), scales=np.zeros((P, 3), np.float32),
             rotations=np.zeros((P, 4), np.float32))
    feat = i["colors_pre"] if i["colors_pre"] is not None else s["rgb"]
    L.orc_render_bwd(P, W, H, _p(s["ranges"]), _p(s["point_list"]), _p(i["bg"]), _p(s["xy"]), _p(s["conic_opacity"]),
                     _p(feat), _p(s["depths"]), _p(s["final_T"]), _p(s["n_contrib"]), _p(dL_dpix), _p(dL_ddepth),
                     _p(g["means2D"]), _p(g["conic"]), _p(g["opacity"]), _p(g["colors"]), _p(g["depths"]))
    fx = W / (2.0 * i["tanx"]); fy = H / (2.0 * i["tany"])
    cov = i["cov3d_pre"] if i["cov3d_pre"] is not None else s["cov3D"]
    L.orc_preprocess_bwd(P, i["D"], M, _p(i["means"]), _p(s["radii"]), _p(i["shs"]) if i["colors_pre"] is None else None,
                         _p(s["clamped"]), _p(i["scales"]), _p(i["rots"]), ctypes.c_float(i["mod"]), _p(cov), _p(i["view"]),
                         _p(i["proj"]), ctypes.c_float(np.float32(fx)), ctypes.c_float(np.float32(fy)),
                         ctypes.c_float(i["tanx"]), ctypes.c_float(i["tany"]), _p(i["campos"]), _p(g["means2D"]),
                         _p(g["conic"]), _p(g["colors"]), _p(g["depths"]), _p(g["means3D"]), _p(g["cov3D"]), _p(g["sh"]),
                         _p(g["scales"]), _p(g["rotations"]))
    return g


def dist2(points):
    pts = _f(points)
    out = np.zeros(pts.shape[0], np.float32)
    lib().orc_dist2(pts.shape[0], _p(pts), _p(out))
    return out


def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-15):
    """In-place Adam on float32 numpy arrays; `step` is the already-incremented step count."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    lib().orc_adam(ctypes.c_longlong(p.size), _p(p), _p(g), _p(m), _p(v), ctypes.c_double(beta1), ctypes.c_double(beta2),
                   ctypes.c_double(eps), ctypes.c_double((lr / bc1) * -1), ctypes.c_double(bc2 ** 0.5))
