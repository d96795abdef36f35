"""Model files of scene/gaussian_model.py for `engine.GaussianState`: the same bytes on disk, so that a
point cloud or a deformation checkpoint written by either side is read by the other.

    save_ply / load_ply                 gaussian_model.py:342-360, 367-407 (+ construct_list_of_attributes :300-312)
    save_deformation / load_model       gaussian_model.py:334-339, 321-333

The PLY is binary little-endian with ONE `vertex` element of 62 float properties at SH degree 3 (x y z, three zero normals, f_dc_*
and f_rest_* CHANNEL-major -- the [P,K,3] tensors transposed to [P,3,K] before flattening --, opacity, scale_*, rot_*).  The
reference fills its structured array through `list(map(tuple, attributes))`, a Python loop over the points (tens of seconds at
1M); here the [P,62] float matrix is viewed as the structured dtype, which is the same memory.  Pure host code: tensors cross
PCIe once each way."""
import os

import numpy as np
import torch
from torch import nn


def ply_attributes(n_dc, n_rest, n_scale=3, n_rot=4):
    """construct_list_of_attributes (gaussian_model.py:300-312)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)] + [f"f_rest_{i}" for i in range(n_rest)] + ["opacity"]
    names += [f"scale_{i}" for i in range(n_scale)] + [f"rot_{i}" for i in range(n_rot)]
    return names


def save_ply(model, path):
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    host = lambda t: t.detach().to("cpu", torch.float32)
    xyz = host(model._xyz)
    f_dc = host(model._features_dc).transpose(1, 2).flatten(start_dim=1)
    f_rest = host(model._features_rest).transpose(1, 2).flatten(start_dim=1)
    cols = torch.cat((xyz, torch.zeros_like(xyz), f_dc, f_rest, host(model._opacity), host(model._scaling), host(model._rotation)),
                     dim=1).contiguous().numpy().astype("<f4", copy=False)
    names = ply_attributes(f_dc.shape[1], f_rest.shape[1], model._scaling.shape[1], model._rotation.shape[1])
    assert cols.shape[1] == len(names)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {cols.shape[0]}"]
    header += [f"property float {n}" for n in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(cols.tobytes())


_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "char": "i1",
              "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2", "int": "<i4", "int32": "<i4",
              "uint": "<u4", "uint32": "<u4"}


def read_ply_vertices(path):
    """The first element of a binary little-endian PLY as a numpy structured array (what `PlyData.read(path).elements[0]` holds)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = f.readline().split()
        if len(fmt) < 2 or fmt[1] != b"binary_little_endian":
            raise ValueError(f"{path}: only binary little-endian PLY is supported")
        elems = []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "end_header":
                break
            if tok[0] == "element":
                elems.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported")
                elems[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
        if not elems:
            raise ValueError(f"{path}: no elements")
        _, count, props = elems[0]
        dt = np.dtype(props)
        raw = f.read(dt.itemsize * count)
        if len(raw) != dt.itemsize * count:
            raise ValueError(f"{path}: truncated data ({len(raw)} of {dt.itemsize * count} bytes)")
        return np.frombuffer(raw, dtype=dt, count=count)


def load_ply(model, path, device=None):
    """gaussian_model.py:367-407: replaces the six per-Gaussian parameters (f_rest_* / scale_* / rot_* sorted by their numeric
    suffix, SH count checked against max_sh_degree) and sets active_sh_degree = max_sh_degree.  Optimiser state is NOT carried over
    (the reference's load_ply does not either): call training_setup again before training."""
    v = read_ply_vertices(path)
    device = device if device is not None else model._xyz.device
    names = v.dtype.names
    col = lambda n: np.asarray(v[n], dtype=np.float32)
    numbered = lambda prefix: sorted((n for n in names if n.startswith(prefix)), key=lambda s: int(s.split("_")[-1]))
    xyz = np.stack((col("x"), col("y"), col("z")), axis=1)
    dc = np.stack((col("f_dc_0"), col("f_dc_1"), col("f_dc_2")), axis=1)[:, :, None]                    # [P,3,1]
    rest_names = numbered("f_rest_")
    if len(rest_names) != 3 * (model.max_sh_degree + 1) ** 2 - 3:
        raise ValueError(f"{path}: {len(rest_names)} f_rest_* properties, SH degree {model.max_sh_degree} needs "
                         f"{3 * (model.max_sh_degree + 1) ** 2 - 3}")
    P = xyz.shape[0]
    rest = (np.stack([col(n) for n in rest_names], axis=1) if rest_names else np.zeros((P, 0), np.float32)).reshape(P, 3, len(rest_names) // 3)
    scales = np.stack([col(n) for n in numbered("scale_")], axis=1)
    rots = np.stack([col(n) for n in numbered("rot")], axis=1)
    opac = col("opacity")[:, None]
    par = lambda a: nn.Parameter(torch.tensor(a, dtype=torch.float32, device=device).requires_grad_(True))
    model._xyz = par(xyz)
    model._features_dc = nn.Parameter(par(dc).data.transpose(1, 2).contiguous().requires_grad_(True))
    model._features_rest = nn.Parameter(par(rest).data.transpose(1, 2).contiguous().requires_grad_(True))
    model._opacity, model._scaling, model._rotation = par(opac), par(scales), par(rots)
    model.active_sh_degree = model.max_sh_degree
    model.optimizer = None
    return model


def save_deformation(model, path):
    """gaussian_model.py:334-339: deformation.pth (the field's state_dict), deformation_table.pth, deformation_accum.pth,
    scene_flow.pth."""
    os.makedirs(path, exist_ok=True)
    P, dev = model._xyz.shape[0], model._xyz.device
    table = getattr(model, "_deformation_table", None)
    accum = getattr(model, "_deformation_accum", None)
    torch.save(model._deformation.state_dict(), os.path.join(path, "deformation.pth"))
    torch.save(table if table is not None else torch.ones(P, dtype=torch.bool, device=dev), os.path.join(path, "deformation_table.pth"))
    torch.save(accum if accum is not None else torch.zeros(P, 3, device=dev), os.path.join(path, "deformation_accum.pth"))
    torch.save(model._scene_flow, os.path.join(path, "scene_flow.pth"))


def load_model(model, path, device=None):
    """gaussian_model.py:321-333."""
    device = device if device is not None else model._xyz.device
    load = lambda name: torch.load(os.path.join(path, name), map_location=device, weights_only=True)
    model._deformation.load_state_dict(load("deformation.pth"))
    model._deformation = model._deformation.to(device)
    flow = load("scene_flow.pth")
    if "_scene_flow" in model._buffers:
        model._buffers["_scene_flow"] = flow.to(device)
    else:
        model._scene_flow = flow.to(device)
    P = model._xyz.shape[0]
    model._deformation_table = torch.ones(P, dtype=torch.bool, device=device)
    model._deformation_accum = torch.zeros(P, 3, device=device)
    for name, attr in (("deformation_table.pth", "_deformation_table"), ("deformation_accum.pth", "_deformation_accum")):
        if os.path.exists(os.path.join(path, name)):
            setattr(model, attr, load(name))
    model.max_radii2D = torch.zeros(P, device=device)
    return model
