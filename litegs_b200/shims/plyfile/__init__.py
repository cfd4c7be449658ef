"""Minimal ``plyfile`` stand-in: the subset the reference uses to write / read point clouds
(litegs/io_manager/ply.py:7-86, litegs/io_manager/colmap.py:281-306).  Numpy structured arrays <-> PLY, binary
little-endian on write (plyfile's default), binary little/big endian and ascii on read.  Scalar properties only."""
from __future__ import annotations

import numpy as np

_PLY_TO_NP = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}
_NP_TO_PLY = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}


class PlyProperty:
    def __init__(self, name: str, val_dtype: str):
        self.name = name
        self.val_dtype = val_dtype

    def __repr__(self):
        return f"PlyProperty({self.name!r}, {self.val_dtype!r})"


class PlyElement:
    def __init__(self, name: str, data: np.ndarray):
        self.name = name
        self.data = data
        self.properties = tuple(PlyProperty(n, data.dtype[n].str.lstrip("<>|=")) for n in data.dtype.names)

    @staticmethod
    def describe(data: np.ndarray, name: str, **_ignored) -> "PlyElement":
        if data.dtype.names is None:
            raise ValueError("PlyElement.describe expects a numpy structured array")
        for n in data.dtype.names:
            if data.dtype[n].shape != () or data.dtype[n].str.lstrip("<>|=") not in _NP_TO_PLY:
                raise ValueError(f"property {n!r}: only scalar numeric properties are supported by this stand-in")
        return PlyElement(name, data)

    @property
    def count(self) -> int:
        return len(self.data)

    def __getitem__(self, key):
        return self.data[key]

    def __len__(self):
        return len(self.data)


class PlyData:
    def __init__(self, elements=(), text: bool = False, byte_order: str = "<", comments=(), obj_info=()):
        self.elements = list(elements)
        self.text = text
        self.byte_order = byte_order
        self.comments = list(comments)

    def __getitem__(self, name: str) -> PlyElement:
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def __contains__(self, name):
        return any(e.name == name for e in self.elements)

    def write(self, stream) -> None:
        own = isinstance(stream, (str, bytes)) or hasattr(stream, "__fspath__")
        f = open(stream, "wb") if own else stream
        try:
            fmt = "ascii" if self.text else ("binary_little_endian" if self.byte_order in ("<", "=") else "binary_big_endian")
            head = ["ply", f"format {fmt} 1.0"] + [f"comment {c}" for c in self.comments]
            for e in self.elements:
                head.append(f"element {e.name} {len(e.data)}")
                for n in e.data.dtype.names:
                    head.append(f"property {_NP_TO_PLY[e.data.dtype[n].str.lstrip('<>|=')]} {n}")
            head.append("end_header")
            f.write(("\n".join(head) + "\n").encode("ascii"))
            for e in self.elements:
                if self.text:
                    for row in e.data:
                        f.write((" ".join(repr(x.item()) for x in row) + "\n").encode("ascii"))
                else:
                    order = "<" if fmt == "binary_little_endian" else ">"
                    dt = np.dtype([(n, order + e.data.dtype[n].str.lstrip("<>|=")) for n in e.data.dtype.names])
                    f.write(np.ascontiguousarray(e.data.astype(dt)).tobytes())
        finally:
            if own:
                f.close()

    @staticmethod
    def read(stream) -> "PlyData":
        own = isinstance(stream, (str, bytes)) or hasattr(stream, "__fspath__")
        f = open(stream, "rb") if own else stream
        try:
            if f.readline().strip() != b"ply":
                raise ValueError("not a PLY file")
            fmt, elements, comments = None, [], []
            while True:
                line = f.readline()
                if not line:
                    raise ValueError("PLY header has no end_header")
                tok = line.decode("ascii").split()
                if not tok:
                    continue
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] == "comment":
                    comments.append(" ".join(tok[1:]))
                elif tok[0] == "element":
                    elements.append((tok[1], int(tok[2]), []))
                elif tok[0] == "property":
                    if tok[1] == "list":
                        raise ValueError("list properties are not supported by this stand-in")
                    elements[-1][2].append((tok[2], _PLY_TO_NP[tok[1]]))
                elif tok[0] == "end_header":
                    break
            out = []
            for name, count, props in elements:
                if fmt == "ascii":
                    dt = np.dtype([(n, "<" + t) for n, t in props])
                    arr = np.empty(count, dtype=dt)
                    for i in range(count):
                        vals = f.readline().split()
                        arr[i] = tuple(np.dtype(t).type(float(v)) if t[0] == "f" else np.dtype(t).type(int(float(v)))
                                       for v, (_, t) in zip(vals, props))
                else:
                    order = "<" if fmt == "binary_little_endian" else ">"
                    dt = np.dtype([(n, order + t) for n, t in props])
                    arr = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
                    arr = arr.astype(np.dtype([(n, "<" + t) for n, t in props]))
                out.append(PlyElement(name, arr))
            return PlyData(out, text=(fmt == "ascii"), byte_order="<" if fmt != "binary_big_endian" else ">", comments=comments)
        finally:
            if own:
                f.close()
