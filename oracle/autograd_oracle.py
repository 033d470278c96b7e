"""PyTorch-autograd restatement of the rasterizer's FORWARD only; gradients come from autograd.   TEST INFRASTRUCTURE ONLY.

Why a second oracle: ``oracle/sagars_oracle.c`` restates the reference's hand-derived backward kernels
(CF cuda_rasterizer/backward.cu) line by line, so a misreading of a formula there would be reproduced faithfully.
Here only the forward expressions are written down -- in float64, as tensor
expressions, tile by tile -- and ``torch.autograd`` derives every gradient independently.  The places where the
reference's backward is NOT the derivative of its forward are modelled explicitly as stop-gradients:

  * A.14  alpha = min(0.99, o*G) is straight-through (CF backward.cu:499,540,556 use ``o * dL_dalpha`` also when the
          clamp was active); the ``power > 0``, ``alpha < 1/255`` and ``T < 1e-4`` tests are constants;
  * A.20  the field-of-view clamp of cov2D: a clamped t.x (t.y) is a constant -- no gradient to t.x and no
          d(clamp)/d(t.z) term (CF backward.cu:175-176, 262-264);
  * A.23  dL_dscales is the gradient w.r.t. ``scale_modifier * scale`` without the factor ``scale_modifier``
          (CF backward.cu:297-325) -- found by this oracle's first scale_modifier != 1 run, not listed in SURVEY.md;
  * DEPTH the mask output only produces dL_dmask (weights detached), the depth output produces no gradient
          (DEPTH backward.cu:516, __init__.py:126-182).

It is also BASELINE.json's configs[0] ("CPU / PyTorch-autograd reference path, no GPU"): ``python -m
oracle.autograd_oracle`` times one forward+backward of SYN(10k, 256x256, K=3) on the host cores.

Only ``tests/`` and ``bench.py``'s cpu legs may import this module; the product never does.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

F64 = torch.float64
TILE = 16

# real spherical-harmonics constants (the published 3DGS basis; same values as CF cuda_rasterizer/auxiliary.h:23-39)
_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def _sh_to_rgb(deg: int, shs: torch.Tensor, means3D: torch.Tensor, campos: torch.Tensor) -> torch.Tensor:
    """CF forward.cu:23-74: colour = max(0, SH(dir) + 0.5), dir = normalise(mean - campos).
    ``clamp_min`` has zero gradient where it clamps, which is what the reference's ``clamped`` flags do (backward.cu:31-34)."""
    d = means3D - campos
    d = d / d.norm(dim=1, keepdim=True)
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    S = lambda i: shs[:, i, :]
    r = _C0 * S(0)
    if deg > 0:
        r = r - _C1 * y * S(1) + _C1 * z * S(2) - _C1 * x * S(3)
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            r = (r + _C2[0] * xy * S(4) + _C2[1] * yz * S(5) + _C2[2] * (2.0 * zz - xx - yy) * S(6) +
                 _C2[3] * xz * S(7) + _C2[4] * (xx - yy) * S(8))
            if deg > 2:
                r = (r + _C3[0] * y * (3.0 * xx - yy) * S(9) + _C3[1] * xy * z * S(10) +
                     _C3[2] * y * (4.0 * zz - xx - yy) * S(11) + _C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * S(12) +
                     _C3[4] * x * (4.0 * zz - xx - yy) * S(13) + _C3[5] * z * (xx - yy) * S(14) +
                     _C3[6] * x * (xx - 3.0 * yy) * S(15))
    return torch.clamp_min(r + 0.5, 0.0)


def _cov3d(scales: torch.Tensor, rotations: torch.Tensor, mod: float) -> torch.Tensor:
    """CF forward.cu:121-155: Sigma = R S^2 R^T, quaternion (r,x,y,z) used as given (not normalised)."""
    r, x, y, z = rotations.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)
    # the reference's dL_dscale is the derivative w.r.t. s = mod * scale, NOT multiplied by mod (CF backward.cu:297-325):
    # value mod * scale, gradient as if d(s)/d(scale) = 1.  Invisible at the scale_modifier = 1 every training script uses.
    s = scales + ((mod - 1.0) * scales).detach()
    M = R * s.unsqueeze(1)                         # R @ diag(s)
    return M @ M.transpose(1, 2)


def _sym6_to_mat(c6: torch.Tensor) -> torch.Tensor:
    a, b, c, d, e, f = c6.unbind(1)
    return torch.stack([a, b, c, b, d, e, c, e, f], dim=1).reshape(-1, 3, 3)


def forward(*, means3D, opacities, bg, viewmatrix, projmatrix, campos, image_height, image_width, tanfovx, tanfovy,
            scale_modifier=1.0, sh_degree=0, shs=None, colors_precomp=None, scales=None, rotations=None,
            cov3D_precomp=None, mask=None, means2D=None):
    """Differentiable forward.  All tensor arguments are torch tensors (any float dtype; computed in float64); those
    with ``requires_grad`` receive gradients.  ``means2D`` [P,3] is the reference's dummy screen-space carrier: it is
    added (as zeros) to the NDC position so that its gradient is the reference's dL_dmean2D (x0.5W, x0.5H, z = 0).
    Returns a namespace: color [C,H,W], out_mask / out_depth [1,H,W] (when ``mask`` is given), final_T, n_contrib,
    radii, tiles_touched, num_rendered, point_list, ranges (integer state as numpy arrays)."""
    H, W = int(image_height), int(image_width)
    P = means3D.shape[0]
    c = lambda t: None if t is None else t.to(F64)
    means3D, opac, view, proj, campos = c(means3D), c(opacities).reshape(-1), c(viewmatrix), c(projmatrix), c(campos).reshape(-1)
    bg = c(bg).reshape(-1)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE

    # ---- per-Gaussian stage (A.1 - A.8) ----
    ones = torch.ones(P, 1, dtype=F64)
    ph = torch.cat([means3D, ones], dim=1)
    t = (ph @ view)[:, :3]                                    # transformPoint4x3 (row-vector convention)
    depth = t[:, 2]
    near_ok = depth.detach() > 0.2
    cov3 = _sym6_to_mat(c(cov3D_precomp)) if cov3D_precomp is not None else _cov3d(c(scales), c(rotations), scale_modifier)
    tz = t[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    rx, ry = (t[:, 0] / tz).detach(), (t[:, 1] / tz).detach()
    tx = torch.where((rx < -limx) | (rx > limx), (rx.clamp(-limx, limx) * tz.detach()), t[:, 0])   # A.20: clamped -> constant
    ty = torch.where((ry < -limy) | (ry > limy), (ry.clamp(-limy, limy) * tz.detach()), t[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * tx / (tz * tz), zero, fy / tz, -fy * ty / (tz * tz), zero, zero, zero], dim=1).reshape(-1, 3, 3)
    Rw = view[:3, :3].transpose(0, 1)                          # world -> view rotation
    Tm = J @ Rw
    cov2 = Tm @ cov3 @ Tm.transpose(1, 2)
    a, b, cc = cov2[:, 0, 0] + 0.3, cov2[:, 0, 1], cov2[:, 1, 1] + 0.3
    det = a * cc - b * b
    det_ok = det.detach() != 0
    det_safe = torch.where(det_ok, det, torch.ones_like(det))
    conic = torch.stack([cc / det_safe, -b / det_safe, a / det_safe], dim=1)
    mid = 0.5 * (a + cc).detach()
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det.detach(), 0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam))
    p_hom = ph @ proj
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w.unsqueeze(1)
    if means2D is not None:
        ndc = ndc + c(means2D)[:, :2]
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    # tile rectangle (auxiliary.h:46-56): fp32 arithmetic and C truncation, as the reference does it
    pxf, pyf, rf = px.detach().to(torch.float32), py.detach().to(torch.float32), radius.to(torch.float32)
    rect = lambda v, g: torch.clamp(torch.trunc(v / 16.0).to(torch.int64), 0, g)
    x0, x1 = rect(pxf - rf, gx), rect(pxf + rf + 15.0, gx)
    y0, y1 = rect(pyf - rf, gy), rect(pyf + rf + 15.0, gy)
    tiles = (x1 - x0) * (y1 - y0)
    vis = near_ok & det_ok & (tiles > 0)
    radii = torch.where(vis, radius.to(torch.int64), torch.zeros_like(tiles))
    tiles = torch.where(vis, tiles, torch.zeros_like(tiles))

    if colors_precomp is not None:
        feat = c(colors_precomp)
    else:
        feat = _sh_to_rgb(int(sh_degree), c(shs), means3D, campos)
    C = feat.shape[1]
    maskv = None if mask is None else c(mask).reshape(-1)

    # ---- binning (A.9): instances ordered by (tile, depth bits, emission order) ----
    depth32 = depth.detach().to(torch.float32).numpy()
    lists = [[] for _ in range(gx * gy)]
    x0n, x1n, y0n, y1n = x0.numpy(), x1.numpy(), y0.numpy(), y1.numpy()
    for i in np.nonzero(vis.numpy())[0]:
        for yy in range(y0n[i], y1n[i]):
            for xx in range(x0n[i], x1n[i]):
                lists[yy * gx + xx].append(i)
    point_list, ranges = [], np.zeros((gx * gy, 2), np.uint32)
    for tid, L in enumerate(lists):
        if L:
            L = np.asarray(L, np.int64)
            L = L[np.argsort(depth32[L], kind="stable")]
            lists[tid] = L
            ranges[tid] = (len(point_list), len(point_list) + len(L))
            point_list.extend(L.tolist())

    # ---- per-tile blend (A.10 - A.12) ----
    color = torch.zeros(C, H, W, dtype=F64) + bg[:C].reshape(C, 1, 1)     # empty tiles: T = 1 -> background
    out_mask = torch.zeros(1, H, W, dtype=F64)
    out_depth = torch.zeros(1, H, W, dtype=F64)
    final_T = torch.ones(H, W, dtype=F64)
    n_contrib = np.zeros((H, W), np.uint32)
    for tid, L in enumerate(lists):
        if len(L) == 0:
            continue
        ty0, tx0 = (tid // gx) * TILE, (tid % gx) * TILE
        ty1, tx1 = min(ty0 + TILE, H), min(tx0 + TILE, W)
        ys, xs = torch.meshgrid(torch.arange(ty0, ty1, dtype=F64), torch.arange(tx0, tx1, dtype=F64), indexing="ij")
        ys, xs = ys.reshape(-1, 1), xs.reshape(-1, 1)
        idx = torch.as_tensor(L)
        dx, dy = px[idx].unsqueeze(0) - xs, py[idx].unsqueeze(0) - ys                       # [npix, L]
        cn = conic[idx]
        power = -0.5 * (cn[:, 0] * dx * dx + cn[:, 2] * dy * dy) - cn[:, 1] * dx * dy
        raw = opac[idx].unsqueeze(0) * torch.exp(power)
        alpha = raw + (torch.clamp_max(raw, 0.99) - raw).detach()                           # A.14 straight-through
        valid = (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
        a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a_eff
        T_before = torch.cumprod(torch.cat([torch.ones_like(one_m[:, :1]), one_m[:, :-1]], dim=1), dim=1)
        stop = valid & ((T_before * one_m).detach() < 1e-4)
        done = torch.cumsum(stop.to(torch.int64), dim=1) > 0                                # the stopping instance is not blended
        contrib = valid & ~done
        w = torch.where(contrib, a_eff * T_before, torch.zeros_like(a_eff))                 # alpha * T
        Tfin = torch.prod(torch.where(contrib, one_m, torch.ones_like(one_m)), dim=1)
        pix = w @ feat[idx] + Tfin.unsqueeze(1) * bg[:C].unsqueeze(0)                       # [npix, C]
        hh, ww = ty1 - ty0, tx1 - tx0
        color[:, ty0:ty1, tx0:tx1] = pix.transpose(0, 1).reshape(C, hh, ww)
        final_T[ty0:ty1, tx0:tx1] = Tfin.detach().reshape(hh, ww)
        pos = torch.arange(1, len(L) + 1).unsqueeze(0)
        n_contrib[ty0:ty1, tx0:tx1] = torch.where(contrib, pos, torch.zeros_like(pos)).max(dim=1).values.reshape(hh, ww).numpy()
        if maskv is not None:
            wd = w.detach()
            out_mask[0, ty0:ty1, tx0:tx1] = (wd @ maskv[idx]).reshape(hh, ww)
            out_depth[0, ty0:ty1, tx0:tx1] = (wd @ depth.detach()[idx]).reshape(hh, ww)
    o = SimpleNamespace(color=color, final_T=final_T.numpy().astype(np.float32), n_contrib=n_contrib,
                        radii=radii.numpy().astype(np.int32), tiles_touched=tiles.numpy().astype(np.uint32),
                        num_rendered=len(point_list), point_list=np.asarray(point_list, np.uint32), ranges=ranges,
                        out_mask=out_mask if maskv is not None else None, out_depth=out_depth if maskv is not None else None)
    return o


def run_scene(sc, K: int, depth: bool = False, use_sh: bool = False, sh_degree: int = 0, bg=None, backward: bool = True):
    """Same contract as tests.common.run_oracle: a namespace of numpy outputs / gradients for a synthetic scene.
    Runs on ONE torch thread: the per-tile tensors are tiny and intra-op threading costs 10x here."""
    nthr = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        return _run_scene(sc, K, depth, use_sh, sh_degree, bg, backward)
    finally:
        torch.set_num_threads(nthr)


def _run_scene(sc, K, depth, use_sh, sh_degree, bg, backward):
    g, cam = sc.gauss, sc.cam
    leaf = lambda t: t.detach().clone().to(F64).requires_grad_(True)
    means3D, opac, scales, rots = leaf(g.means3D), leaf(g.opacities), leaf(g.scales), leaf(g.rotations)
    means2D = torch.zeros(sc.P, 3, dtype=F64, requires_grad=True)
    colors = None if use_sh else leaf(g.colors)
    shs = leaf(g.shs) if use_sh else None
    mask = None
    if depth:
        mask = leaf(torch.rand(sc.P, 1, generator=torch.Generator().manual_seed(7)) * 0.5 + 0.5)
    bg_t = torch.zeros(max(K, 3)) if bg is None else bg
    fw = forward(means3D=means3D, opacities=opac, bg=bg_t, viewmatrix=cam.world_view_transform,
                 projmatrix=cam.full_proj_transform, campos=cam.camera_center, image_height=sc.H, image_width=sc.W,
                 tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, sh_degree=sh_degree, shs=shs, colors_precomp=colors, scales=scales,
                 rotations=rots, mask=mask, means2D=means2D)
    o = SimpleNamespace(kind="autograd", variant="depth" if depth else ("base" if K == 3 else "cf"))
    o.color = fw.color.detach().numpy().astype(np.float32)
    o.out_mask = None if not depth else fw.out_mask.detach().numpy().astype(np.float32)
    o.out_depth = None if not depth else fw.out_depth.detach().numpy().astype(np.float32)
    for k in ("final_T", "n_contrib", "radii", "tiles_touched", "num_rendered", "point_list", "ranges"):
        setattr(o, k, getattr(fw, k))
    o.point_offsets = np.cumsum(fw.tiles_touched.astype(np.uint64)).astype(np.uint32)
    if backward:
        loss = (fw.color * sc.dL_dout[:K].to(F64)).sum()
        if depth:
            loss = loss + (fw.out_mask * sc.dL_dmask.to(F64)).sum()
        loss.backward()
        gr = lambda t: None if t is None or t.grad is None else t.grad.numpy().astype(np.float32)
        o.g_means3D, o.g_means2D, o.g_opacity = gr(means3D), gr(means2D), gr(opac)
        o.g_scales, o.g_rotations = gr(scales), gr(rots)
        o.g_colors, o.g_sh = gr(colors), gr(shs)
        o.g_mask = gr(mask).reshape(-1) if depth else None
        o.g_cov3D = None
    return o


if __name__ == "__main__":   # BASELINE.json configs[0] on the host cores
    import os
    import sys
    import time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from seganygaussians_b200 import synthetic
    P, H, W, K = 10_000, 256, 256, 3
    sc = synthetic.scene(P, H, W, K)
    t0 = time.time()
    out = run_scene(sc, K)
    dt = time.time() - t0
    print(f"c1 SYN({P}, {H}x{W}, K={K}): R={out.num_rendered} fwd+bwd {dt:.2f} s on 1 thread "
          f"-> {P * H * W / dt:.3e} Gaussian*pixel/s (PyTorch autograd, float64)")
