"""The per-Gaussian kernels of the product (csrc/preprocess_kernels.cuh: cull, cov3D, EWA cov2D, conic, radius, tile rectangle,
SH -> RGB, packed record; csrc/geom_backward_kernels.cuh: conic -> cov2D -> cov3D -> scale / rotation, mean2D -> mean3D, SH
backward) executed on the CPU under the CUDA execution shim and compared with the CPU oracle.

The host compiler does not contract a*b+c the way nvcc does (the oracle models nvcc's contraction with explicit fmaf), so
floats agree to fp32 rounding rather than bit for bit, and a radius that sits within an ulp of an integer may flip: integer
state must agree for all but at most 0.2 % of the Gaussians; on the GPU it is bit-exact (tests/test_parity_gpu.py)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from tests import common
from seganygaussians_b200 import synthetic
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32p = C.c_void_p


@pytest.fixture(scope="module")
def emu():
    d = tempfile.mkdtemp(prefix="sagars_emu_")
    so = os.path.join(d, "libemu_geometry.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "tests", "cuda_emu"), "-I", os.path.join(ROOT, "seganygaussians_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cuda_emu", "emu_geometry.cpp"), "-o", so])
    L = C.CDLL(so)
    L.emu_geom_bytes.restype = C.c_size_t
    L.emu_geom_bytes.argtypes = [C.c_int]
    L.emu_geom_offsets.argtypes = [C.c_int, C.c_void_p]
    L.emu_preprocess.argtypes = [C.c_int] * 4 + [f32p, f32p, C.c_float, f32p, f32p, f32p, f32p, C.c_int, f32p, f32p, f32p, C.c_int, C.c_int,
                                                 C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_uint]
    L.emu_geom_backward.argtypes = [C.c_int] * 3 + [f32p] * 7 + [C.c_float] + [f32p] * 3 + [C.c_int, C.c_int, C.c_float, C.c_float] + [f32p] * 10
    return L


def _p(a):
    return None if a is None else a.ctypes.data


def _c(t):
    return None if t is None else np.ascontiguousarray(t.numpy().astype(np.float32))


CASES = [("precomp_k3", 1500, 60, 80, False, 0, False, 2.0, 8.0), ("sh3", 1200, 48, 64, True, 3, False, 3.0, 8.0),
         ("sh1_close_camera", 900, 48, 64, True, 1, False, 6.0, 3.0),      # near-plane culls, field-of-view clamps
         ("cov3d_precomp", 800, 40, 56, False, 0, True, 3.0, 8.0)]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_preprocess_and_geometry_backward(emu, case):
    name, P, H, W, use_sh, deg, cov_pre, sigma, cam_radius = case
    K = 3
    sc = synthetic.scene(P, H, W, K, sh_coeffs=16 if use_sh else 0, sigma_px=sigma)
    sc.cam = synthetic.make_camera(H, W, 0, radius=cam_radius)
    g, c = sc.gauss, sc.cam
    cov = None
    if cov_pre:
        from oracle import autograd_oracle as ag
        S = ag._cov3d(g.scales.to(torch.float64), g.rotations.to(torch.float64), 1.0)
        cov = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=1).to(torch.float32).contiguous()
    kw = dict(means3D=g.means3D.numpy(), opacities=g.opacities.numpy(), bg=np.zeros(3, np.float32), viewmatrix=c.world_view_transform.numpy(),
              projmatrix=c.full_proj_transform.numpy(), campos=c.camera_center.numpy(), image_height=H, image_width=W, tanfovx=c.tanfovx,
              tanfovy=c.tanfovy, sh_degree=deg, shs=g.shs.numpy() if use_sh else None, colors_precomp=None if use_sh else g.colors.numpy(),
              scales=None if cov_pre else g.scales.numpy(), rotations=None if cov_pre else g.rotations.numpy(),
              cov3D_precomp=None if cov is None else cov.numpy(), num_channels=K)
    fw = orc.forward(**kw)
    bw = orc.backward(fw, sc.dL_dout[:K].numpy())

    # ---- forward per-Gaussian stage ----
    means3D, scales, rots, opac = _c(g.means3D), (None if cov_pre else _c(g.scales)), (None if cov_pre else _c(g.rotations)), _c(g.opacities)
    shs = _c(g.shs) if use_sh else None
    covp = _c(cov) if cov_pre else None
    view, proj, campos = _c(c.world_view_transform), _c(c.full_proj_transform), _c(c.camera_center)
    buf = np.zeros(emu.emu_geom_bytes(P) // 4 + 64, np.uint32)      # 4-byte aligned storage, viewed as bytes below
    raw = buf.view(np.uint8)
    radii = np.zeros(P, np.int32)
    emu.emu_preprocess(P, deg, 16 if use_sh else 0, K, _p(means3D), _p(scales), 1.0, _p(rots), _p(opac), _p(shs), _p(covp), 0 if use_sh else 1,
                       _p(view), _p(proj), _p(campos), W, H, c.tanfovx, c.tanfovy, _p(radii), _p(buf), 0)
    off = np.zeros(8, np.uint64)
    emu.emu_geom_offsets(P, _p(off))
    off = [int(x) for x in off]
    view_as = lambda o, dt, n: raw[o:o + n * np.dtype(dt).itemsize].view(dt)
    depths = view_as(off[0], np.float32, P)
    geo = view_as(off[1], np.float32, 8 * P).reshape(P, 8)
    cov3D = view_as(off[2], np.float32, 6 * P).reshape(P, 6)
    rgb = view_as(off[3], np.float32, 3 * P).reshape(P, 3)
    clamped = view_as(off[4], np.uint8, 3 * P).reshape(P, 3)
    tiles = view_as(off[5], np.uint32, P)
    block_sums = view_as(off[6], np.uint32, (P + 255) // 256)
    agree = (radii == fw.radii) & (tiles == fw.tiles_touched)
    assert agree.mean() >= 0.998, f"{(~agree).sum()} of {P} radii / tile counts differ"
    assert (fw.radii > 0).sum() > 0.3 * P
    vis = agree & (fw.radii > 0)

    def close(got, want, what, scale=1.0):
        r, d, s = common.float_err(got, want)
        assert r <= scale, f"{what}: max|d|={d:.3e} max|ref|={s:.3e} tol-ratio={r:.2f}"
    close(geo[vis, 0:2], fw.means2D[vis], "means2D")
    close(geo[vis, 2:5] , fw.conic_opacity[vis, 0:3], "conic", scale=3.0)     # 1 / det of a nearly singular cov2D
    close(geo[vis, 5], fw.conic_opacity[vis, 3], "opacity")
    close(depths[vis], fw.depths[vis], "depths")
    if not cov_pre:
        close(cov3D[vis], fw.cov3D[vis], "cov3D")
    if use_sh:
        close(rgb[vis], fw.rgb[vis], "rgb")
        assert (clamped[vis] != fw.clamped[vis]).mean() < 0.002
    per_block = np.add.reduceat(np.concatenate([tiles, np.zeros((-P) % 256, np.uint32)]).astype(np.uint64), np.arange(0, P + (-P) % 256, 256))
    assert np.array_equal(block_sums.astype(np.uint64), per_block)       # the scan's input, produced in the same pass

    # ---- backward per-Gaussian stage, driven with the oracle's blend-stage accumulators ----
    ggrad = np.zeros((P, 8), np.float32)
    ggrad[:, 0:2] = bw.means2D[:, 0:2]
    ggrad[:, 2:5] = bw.conic[:, [0, 1, 3]]
    ggrad[:, 5] = bw.opacity.reshape(-1)
    dcol = np.ascontiguousarray(bw.colors.astype(np.float32))
    out = {k: np.zeros(s, np.float32) for k, s in dict(m2=(P, 3), op=(P, 1), m3=(P, 3), cov=(P, 6), sh=(P, 16, 3), sc=(P, 3), rot=(P, 4)).items()}
    cov_in = np.ascontiguousarray((cov.numpy() if cov_pre else fw.cov3D).astype(np.float32))
    emu.emu_geom_backward(P, deg, 16 if use_sh else 0, _p(means3D), _p(np.ascontiguousarray(fw.radii)), _p(cov_in), _p(shs),
                          _p(np.ascontiguousarray(fw.clamped)), _p(scales), _p(rots), 1.0, _p(view), _p(proj), _p(campos), W, H, c.tanfovx, c.tanfovy,
                          _p(ggrad), _p(dcol), _p(out["m2"]), _p(out["op"]), None, _p(out["m3"]), _p(out["cov"]), _p(out["sh"]),
                          _p(out["sc"]), _p(out["rot"]))
    close(out["m2"], bw.means2D, "dL_dmeans2D")
    close(out["op"], bw.opacity, "dL_dopacity")
    close(out["m3"], bw.means3D, "dL_dmeans3D", scale=3.0)      # cancellation for near-plane splats (see test_autograd_oracle)
    close(out["cov"], bw.cov3D, "dL_dcov3D")
    if use_sh:
        close(out["sh"], bw.sh, "dL_dsh")
    if not cov_pre:
        close(out["sc"], bw.scales, "dL_dscales")
        close(out["rot"], bw.rotations, "dL_drotations")
