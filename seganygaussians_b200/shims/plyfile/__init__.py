"""Import-compatible stand-in for the ``plyfile`` package, limited to what the SAGA / 3DGS scripts use
(SURVEY.md §8(f) rank 1): the unmodified ``scene/gaussian_model.py``, ``scene/gaussian_model_ff.py`` and
``scene/dataset_readers.py`` do

    from plyfile import PlyData, PlyElement
    plydata = PlyData.read(path); plydata.elements[0]["x"]; plydata['vertex']; [p.name for p in el.properties]
    el = PlyElement.describe(structured_array, 'vertex'); PlyData([el]).write(path)

(``gaussian_model_ff.py:552-592,603-686``, ``gaussian_model.py:213-306``, ``dataset_readers.py:122-147``).
Files are byte-compatible with the real package for scalar properties: header in the same order and spelling,
``binary_little_endian`` payload = the structured array's bytes.  List properties (mesh faces) are read, not written.
This is host-side file I/O only; nothing here touches the GPU path.
"""
from __future__ import annotations

import sys
from typing import Iterable, List

import numpy as np

# PLY scalar type names (both spellings) <-> numpy codes
_PLY_TO_NP = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}
_NP_TO_PLY = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}


class PlyParseError(Exception):
    pass


class PlyProperty:
    def __init__(self, name: str, val_dtype: str):
        self.name = name
        self.val_dtype = val_dtype          # numpy code without byte order, e.g. 'f4'

    def __repr__(self):
        return f"PlyProperty({self.name!r}, {_NP_TO_PLY[self.val_dtype]!r})"

    def header_line(self) -> str:
        return f"property {_NP_TO_PLY[self.val_dtype]} {self.name}"


class PlyListProperty(PlyProperty):
    def __init__(self, name: str, len_dtype: str, val_dtype: str):
        super().__init__(name, val_dtype)
        self.len_dtype = len_dtype

    def header_line(self) -> str:
        return f"property list {_NP_TO_PLY[self.len_dtype]} {_NP_TO_PLY[self.val_dtype]} {self.name}"


class PlyElement:
    def __init__(self, name: str, properties: List[PlyProperty], count: int, data=None):
        self.name = name
        self.properties = tuple(properties)
        self.count = int(count)
        self.data = data

    @staticmethod
    def describe(data: np.ndarray, name: str, len_types=None, val_types=None, comments=None) -> "PlyElement":
        """Element from a 1-D structured numpy array (scalar fields only), like ``plyfile.PlyElement.describe``."""
        if not isinstance(data, np.ndarray) or data.ndim != 1 or data.dtype.names is None:
            raise TypeError("only one-dimensional structured arrays are supported")
        props = []
        for fname in data.dtype.names:
            dt = data.dtype.fields[fname][0]
            if dt.shape != () or dt.kind == "O":
                raise ValueError("list / sub-array properties cannot be written by this stand-in")
            code = dt.str[1:]
            if code not in _NP_TO_PLY:
                raise ValueError(f"unsupported field type {dt} for property {fname!r}")
            props.append(PlyProperty(fname, code))
        return PlyElement(name, props, len(data), data)

    def __getitem__(self, key):
        return self.data[key]

    def __setitem__(self, key, value):
        self.data[key] = value

    def __len__(self):
        return self.count

    def ply_property(self, name: str) -> PlyProperty:
        for p in self.properties:
            if p.name == name:
                return p
        raise KeyError(name)

    def header(self) -> str:
        return "\n".join([f"element {self.name} {self.count}"] + [p.header_line() for p in self.properties])

    def _dtype(self, byte_order: str) -> np.dtype:
        return np.dtype([(p.name, byte_order + p.val_dtype) for p in self.properties])

    @property
    def _has_lists(self) -> bool:
        return any(isinstance(p, PlyListProperty) for p in self.properties)


class PlyData:
    def __init__(self, elements: Iterable[PlyElement] = (), text: bool = False, byte_order: str = "=", comments=None,
                 obj_info=None):
        self.elements = list(elements)
        self.text = bool(text)
        if byte_order == "=":
            byte_order = "<" if sys.byteorder == "little" else ">"
        self.byte_order = byte_order
        self.comments = list(comments or [])
        self.obj_info = list(obj_info or [])

    def __getitem__(self, name: str) -> PlyElement:
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def __contains__(self, name: str) -> bool:
        return any(e.name == name for e in self.elements)

    def __iter__(self):
        return iter(self.elements)

    def __len__(self):
        return len(self.elements)

    # ------------------------------------------------------------------ writing
    @property
    def header(self) -> str:
        fmt = "ascii" if self.text else ("binary_little_endian" if self.byte_order == "<" else "binary_big_endian")
        lines = ["ply", f"format {fmt} 1.0"]
        lines += [f"comment {c}" for c in self.comments]
        lines += [f"obj_info {c}" for c in self.obj_info]
        lines += [e.header() for e in self.elements]
        lines.append("end_header")
        return "\n".join(lines)

    def write(self, stream) -> None:
        own = isinstance(stream, (str, bytes)) or hasattr(stream, "__fspath__")
        f = open(stream, "wb") if own else stream
        try:
            f.write((self.header + "\n").encode("ascii"))
            for e in self.elements:
                if e._has_lists:
                    raise ValueError("list properties cannot be written by this stand-in")
                if self.text:
                    names = [p.name for p in e.properties]
                    for row in e.data:
                        f.write((" ".join(repr(row[n].item()) if e.ply_property(n).val_dtype[0] == "f" else str(row[n].item())
                                          for n in names) + "\n").encode("ascii"))
                else:
                    f.write(np.ascontiguousarray(e.data.astype(e._dtype(self.byte_order), copy=False)).tobytes())
        finally:
            if own:
                f.close()

    # ------------------------------------------------------------------ reading
    @staticmethod
    def read(stream, mmap: bool = False) -> "PlyData":
        own = isinstance(stream, (str, bytes)) or hasattr(stream, "__fspath__")
        f = open(stream, "rb") if own else stream
        try:
            return PlyData._read(f)
        finally:
            if own:
                f.close()

    @staticmethod
    def _read(f) -> "PlyData":
        def line():
            raw = f.readline()
            if not raw:
                raise PlyParseError("unexpected end of header")
            return raw.decode("ascii", errors="replace").strip()

        if line() != "ply":
            raise PlyParseError("not a PLY file (missing 'ply' magic)")
        fmt, comments, obj_info, elements = None, [], [], []
        while True:
            ln = line()
            if not ln:
                continue
            tok = ln.split()
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "comment":
                comments.append(ln[len("comment"):].strip())
            elif tok[0] == "obj_info":
                obj_info.append(ln[len("obj_info"):].strip())
            elif tok[0] == "element":
                elements.append(PlyElement(tok[1], [], int(tok[2])))
            elif tok[0] == "property":
                if not elements:
                    raise PlyParseError("property before any element")
                e = elements[-1]
                if tok[1] == "list":
                    prop = PlyListProperty(tok[4], _PLY_TO_NP[tok[2]], _PLY_TO_NP[tok[3]])
                else:
                    if tok[1] not in _PLY_TO_NP:
                        raise PlyParseError(f"unknown property type {tok[1]!r}")
                    prop = PlyProperty(tok[2], _PLY_TO_NP[tok[1]])
                e.properties = e.properties + (prop,)
            elif tok[0] == "end_header":
                break
            else:
                raise PlyParseError(f"unexpected header line {ln!r}")
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise PlyParseError(f"unsupported format {fmt!r}")
        text = fmt == "ascii"
        bo = ">" if fmt == "binary_big_endian" else "<"
        for e in elements:
            if text:
                e.data = PlyData._read_text(f, e)
            elif e._has_lists:
                e.data = PlyData._read_binary_lists(f, e, bo)
            else:
                dt = e._dtype(bo)
                buf = f.read(dt.itemsize * e.count)
                if len(buf) != dt.itemsize * e.count:
                    raise PlyParseError(f"element {e.name!r}: file is truncated")
                e.data = np.frombuffer(buf, dtype=dt, count=e.count).copy()
        return PlyData(elements, text=text, byte_order=bo, comments=comments, obj_info=obj_info)

    @staticmethod
    def _read_text(f, e: PlyElement) -> np.ndarray:
        if not e._has_lists:
            out = np.empty(e.count, dtype=e._dtype("="))
            names = [p.name for p in e.properties]
            for i in range(e.count):
                tok = f.readline().split()
                if len(tok) < len(names):
                    raise PlyParseError(f"element {e.name!r}: row {i} is short")
                for n, t in zip(names, tok):
                    out[n][i] = float(t) if out.dtype[n].kind == "f" else int(t)
            return out
        out = np.empty(e.count, dtype=[(p.name, "O" if isinstance(p, PlyListProperty) else "=" + p.val_dtype) for p in e.properties])
        for i in range(e.count):
            tok = f.readline().split()
            k = 0
            for p in e.properties:
                if isinstance(p, PlyListProperty):
                    n = int(tok[k]); k += 1
                    out[p.name][i] = np.array(tok[k:k + n], dtype="=" + p.val_dtype); k += n
                else:
                    out[p.name][i] = float(tok[k]) if p.val_dtype[0] == "f" else int(tok[k]); k += 1
        return out

    @staticmethod
    def _read_binary_lists(f, e: PlyElement, bo: str) -> np.ndarray:
        out = np.empty(e.count, dtype=[(p.name, "O" if isinstance(p, PlyListProperty) else bo + p.val_dtype) for p in e.properties])
        for i in range(e.count):
            for p in e.properties:
                if isinstance(p, PlyListProperty):
                    n = int(np.frombuffer(f.read(np.dtype(p.len_dtype).itemsize), dtype=bo + p.len_dtype)[0])
                    out[p.name][i] = np.frombuffer(f.read(np.dtype(p.val_dtype).itemsize * n), dtype=bo + p.val_dtype).copy()
                else:
                    out[p.name][i] = np.frombuffer(f.read(np.dtype(p.val_dtype).itemsize), dtype=bo + p.val_dtype)[0]
        return out


__all__ = ["PlyData", "PlyElement", "PlyProperty", "PlyListProperty", "PlyParseError"]
