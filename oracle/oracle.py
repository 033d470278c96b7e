"""numpy front end of the CPU oracle (``oracle/sagars_oracle.c``).   TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product path (``seganygaussians_b200``) never does.

``forward(...)`` / ``backward(...)`` take and return plain numpy arrays with the reference's shapes, and
also expose every integer intermediate the parity tests compare bit-exactly (radii, tiles_touched,
point_offsets, point_list, sorted keys, ranges, n_contrib).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sagars_oracle.c")
BUILD_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off -fopenmp (explicit fmaf() in the source models nvcc's contraction)."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
        subprocess.check_call(cmd)
    return LIB


class _In(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("C", C.c_int), ("W", C.c_int), ("H", C.c_int),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
                ("has_mask_depth", C.c_int)] + \
               [(n, C.c_void_p) for n in ("bg", "means3D", "shs", "colors_precomp", "opacities", "mask", "scales",
                                          "rotations", "cov3D_precomp", "view", "proj", "campos")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            if os.path.exists(SRC):
                build()
            else:  # pragma: no cover
                raise RuntimeError("oracle library missing")
        _lib = C.CDLL(LIB)
        _lib.orc_preprocess.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _f(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _make_in(keep, *, P, D, M, Cc, W, H, tanfovx, tanfovy, scale_modifier, has_md, bg, means3D, shs, colors, opacities,
             mask, scales, rotations, cov3D_precomp, view, proj, campos):
    s = _In()
    s.P, s.D, s.M, s.C, s.W, s.H = P, D, M, Cc, W, H
    s.tan_fovx, s.tan_fovy, s.scale_modifier = tanfovx, tanfovy, scale_modifier
    s.has_mask_depth = 1 if has_md else 0
    arrs = dict(bg=bg, means3D=means3D, shs=shs, colors_precomp=colors, opacities=opacities, mask=mask, scales=scales,
                rotations=rotations, cov3D_precomp=cov3D_precomp, view=view, proj=proj, campos=campos)
    for k, v in arrs.items():
        v = _f(v)
        keep.append(v)
        setattr(s, k, None if v is None else v.ctypes.data)
    return s


def forward(*, means3D, opacities, bg, viewmatrix, projmatrix, campos, image_height, image_width, tanfovx, tanfovy,
            scale_modifier=1.0, sh_degree=0, shs=None, colors_precomp=None, scales=None, rotations=None,
            cov3D_precomp=None, mask=None, num_channels=None, nthreads=1):
    """CPU forward. Returns a namespace with the outputs and every intermediate buffer."""
    L = lib()
    means3D = _f(means3D)
    P = means3D.shape[0]
    H, W = int(image_height), int(image_width)
    colors_precomp = None if colors_precomp is None or np.size(colors_precomp) == 0 else _f(colors_precomp)
    shs = None if shs is None or np.size(shs) == 0 else _f(shs)
    Cc = int(colors_precomp.shape[1]) if colors_precomp is not None else (num_channels or 3)
    M = int(shs.shape[1]) if shs is not None else 0
    has_md = mask is not None
    keep = []
    s = _make_in(keep, P=P, D=int(sh_degree), M=M, Cc=Cc, W=W, H=H, tanfovx=float(tanfovx), tanfovy=float(tanfovy),
                 scale_modifier=float(scale_modifier), has_md=has_md, bg=np.asarray(bg, dtype=np.float32).reshape(-1),
                 means3D=means3D, shs=shs, colors=colors_precomp, opacities=_f(opacities, (-1,)),
                 mask=None if mask is None else _f(mask, (-1,)),
                 scales=None if scales is None or np.size(scales) == 0 else scales,
                 rotations=None if rotations is None or np.size(rotations) == 0 else rotations,
                 cov3D_precomp=None if cov3D_precomp is None or np.size(cov3D_precomp) == 0 else cov3D_precomp,
                 view=_f(viewmatrix, (-1,)), proj=_f(projmatrix, (-1,)), campos=_f(campos, (-1,)))
    o = SimpleNamespace()
    o.P, o.C, o.M, o.H, o.W, o.has_md = P, Cc, M, H, W, has_md
    o.radii = np.zeros(P, np.int32)
    o.means2D = np.zeros((P, 2), np.float32)
    o.depths = np.zeros(P, np.float32)
    o.cov3D = np.zeros((P, 6), np.float32)
    o.conic_opacity = np.zeros((P, 4), np.float32)
    o.rgb = np.zeros((P, 3), np.float32)
    o.clamped = np.zeros((P, 3), np.uint8)
    o.tiles_touched = np.zeros(P, np.uint32)
    o.point_offsets = np.zeros(P, np.uint32)
    R = L.orc_preprocess(C.byref(s), _p(o.radii), _p(o.means2D), _p(o.depths), _p(o.cov3D), _p(o.conic_opacity),
                         _p(o.rgb), _p(o.clamped), _p(o.tiles_touched), _p(o.point_offsets)) if P > 0 else 0
    o.num_rendered = int(R)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    o.keys = np.zeros(max(R, 1), np.uint64)[:R]
    o.point_list = np.zeros(max(R, 1), np.uint32)[:R]
    o.ranges = np.zeros((gx * gy, 2), np.uint32)
    if P > 0:
        keys = np.zeros(max(R, 1), np.uint64)
        pl = np.zeros(max(R, 1), np.uint32)
        L.orc_bin(C.byref(s), _p(o.radii), _p(o.means2D), _p(o.depths), _p(o.point_offsets), C.c_int(R), _p(keys), _p(pl),
                  _p(o.ranges))
        o.keys, o.point_list = keys[:R], pl[:R]
        o._keys_full, o._pl_full = keys, pl
    o.final_T = np.ones((H, W), np.float32)
    o.n_contrib = np.zeros((H, W), np.uint32)
    o.color = np.zeros((Cc, H, W), np.float32)
    o.out_mask = np.zeros((1, H, W), np.float32)
    o.out_depth = np.zeros((1, H, W), np.float32)
    o.features = colors_precomp if colors_precomp is not None else o.rgb
    if P > 0:
        pl = o._pl_full
        L.orc_render_forward(C.byref(s), _p(o.ranges), _p(pl), _p(o.means2D), _p(o.conic_opacity), _p(o.features),
                             _p(o.depths), _p(o.final_T), _p(o.n_contrib), _p(o.color), _p(o.out_mask), _p(o.out_depth),
                             C.c_int(int(nthreads)))
    else:
        o.color[:] = np.asarray(bg, np.float32).reshape(-1)[:Cc, None, None] * 0  # reference returns zeros for P == 0
    o._in, o._keep = s, keep
    return o


def backward(fw, dL_dout_color, dL_dout_mask=None, nthreads=1):
    """CPU backward for a forward namespace. Returns a namespace of gradients with the reference's shapes."""
    L = lib()
    s = fw._in
    P, Cc, M = fw.P, fw.C, fw.M
    g = SimpleNamespace()
    g.means2D = np.zeros((P, 3), np.float32)
    g.conic = np.zeros((P, 4), np.float32)
    g.opacity = np.zeros((P, 1), np.float32)
    g.colors = np.zeros((P, Cc), np.float32)
    g.mask = np.zeros((P, 1), np.float32)
    g.means3D = np.zeros((P, 3), np.float32)
    g.cov3D = np.zeros((P, 6), np.float32)
    g.sh = np.zeros((P, M, 3), np.float32)
    g.scales = np.zeros((P, 3), np.float32)
    g.rotations = np.zeros((P, 4), np.float32)
    if P == 0:
        return g
    dpix = _f(dL_dout_color)
    dmask = _f(dL_dout_mask) if dL_dout_mask is not None else np.zeros((1, fw.H, fw.W), np.float32)
    L.orc_render_backward(C.byref(s), _p(fw.ranges), _p(fw._pl_full), _p(fw.means2D), _p(fw.conic_opacity), _p(fw.features),
                          _p(fw.final_T), _p(fw.n_contrib), _p(dpix), _p(dmask), _p(g.means2D), _p(g.conic), _p(g.opacity),
                          _p(g.colors), _p(g.mask), C.c_int(int(nthreads)))
    cov3D = fw.cov3D
    if s.cov3D_precomp:
        cov3D = np.ctypeslib.as_array(C.cast(s.cov3D_precomp, C.POINTER(C.c_float)), shape=(P, 6))
    cov3D = np.ascontiguousarray(cov3D)
    L.orc_geom_backward(C.byref(s), _p(fw.radii), _p(cov3D), _p(fw.clamped), _p(g.means2D), _p(g.conic), _p(g.colors),
                        _p(g.means3D), _p(g.cov3D), _p(g.sh), _p(g.scales), _p(g.rotations))
    return g


def mark_visible(means3D, viewmatrix):
    L = lib()
    m = _f(means3D)
    v = _f(viewmatrix, (-1,))
    out = np.zeros(m.shape[0], np.uint8)
    L.orc_mark_visible(C.c_int(m.shape[0]), _p(m), _p(v), _p(out))
    return out.astype(bool)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
