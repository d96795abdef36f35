"""Minimal `plyfile` (binary little-endian, a single `vertex` element), enough for
GaussianModel.save_ply / load_ply (scene/gaussian_model.py:342-407) and storePly/fetchPly."""
import numpy as np

_TYPES = {"f4": "float", "f8": "double", "u1": "uchar", "i4": "int", "u4": "uint", "i2": "short", "u2": "ushort", "i1": "char"}
_REV = {v: k for k, v in _TYPES.items()}
_REV.update({"float32": "f4", "float64": "f8", "uint8": "u1", "int32": "i4"})


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data

    @staticmethod
    def describe(data, name):
        return PlyElement(name, data)

    def __getitem__(self, k):
        return self.data[k]

    @property
    def properties(self):
        return [type("P", (), {"name": n}) for n in self.data.dtype.names]


class PlyData:
    def __init__(self, elements):
        self.elements = list(elements)

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def write(self, path):
        with open(path, "wb") as f:
            hdr = ["ply", "format binary_little_endian 1.0"]
            for e in self.elements:
                hdr.append(f"element {e.name} {len(e.data)}")
                for n in e.data.dtype.names:
                    hdr.append(f"property {_TYPES[e.data.dtype[n].str[1:]]} {n}")
            hdr.append("end_header")
            f.write(("\n".join(hdr) + "\n").encode("ascii"))
            for e in self.elements:
                f.write(np.ascontiguousarray(e.data).astype(e.data.dtype.newbyteorder("<")).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            assert f.readline().strip() == b"ply"
            fmt = f.readline().split()
            assert fmt[1] == b"binary_little_endian", "only binary little-endian PLY is supported"
            elems, cur = [], None
            while True:
                line = f.readline().decode("ascii").split()
                if line[0] == "end_header":
                    break
                if line[0] == "element":
                    cur = [line[1], int(line[2]), []]
                    elems.append(cur)
                elif line[0] == "property":
                    cur[2].append((line[2], "<" + _REV[line[1]]))
            out = []
            for name, count, props in elems:
                dt = np.dtype(props)
                out.append(PlyElement(name, np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)))
        return PlyData(out)
