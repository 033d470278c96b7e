"""Synthetic workloads ``SYN(P, H, W, K, cam)`` of BASELINE.md section 2.2 / SURVEY.md section 8(d).

Generated on the CPU with a seeded ``torch.Generator`` in fp32, so the CPU oracle, the reference CUDA
extension and this library all see identical bits.  Cameras are built exactly like the reference's
``scene/cameras.py:62-65`` with ``utils/graphics_utils.py:38-98`` (world->view transposed, full
projection = view @ proj, camera centre = inverse(view)[3, :3]).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch


def _world2view(R_c2w: np.ndarray, t: np.ndarray) -> np.ndarray:
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R_c2w.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    return np.float32(Rt)


def _projection(znear, zfar, fovx, fovy) -> torch.Tensor:
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    bottom, left = -top, -right
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (right - left)
    Pm[1, 1] = 2.0 * znear / (top - bottom)
    Pm[0, 2] = (right + left) / (right - left)
    Pm[1, 2] = (top + bottom) / (top - bottom)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def make_camera(H: int, W: int, cam: int = 0, fovx: float = 1.0, radius: float = 8.0, n_cams: int = 8):
    """Camera `cam` of `n_cams` on a circle of `radius` in the xz-plane, looking at the origin, y-up."""
    ang = cam * (2.0 * math.pi / n_cams)
    C = np.array([radius * math.sin(ang), 0.0, -radius * math.cos(ang)])
    f = -C / np.linalg.norm(C)
    up = np.array([0.0, 1.0, 0.0])
    x = np.cross(up, f)
    x /= np.linalg.norm(x)
    y = np.cross(f, x)
    R_c2w = np.stack([x, y, f], axis=1)
    t = -R_c2w.T @ C
    tanfovx = math.tan(0.5 * fovx)
    tanfovy = tanfovx * H / W
    fovy = 2.0 * math.atan(tanfovy)
    view = torch.tensor(_world2view(R_c2w, t)).transpose(0, 1).contiguous()
    proj = _projection(0.01, 100.0, fovx, fovy).transpose(0, 1)
    full = (view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    campos = view.inverse()[3, :3].contiguous()
    return SimpleNamespace(image_height=H, image_width=W, feature_height=H, feature_width=W, FoVx=fovx, FoVy=fovy,
                           tanfovx=tanfovx, tanfovy=tanfovy, world_view_transform=view, full_proj_transform=full,
                           camera_center=campos, znear=0.01, zfar=100.0)


def make_gaussians(P: int, K: int, W: int, fovx: float = 1.0, seed: int = 0, sigma_px: float = 2.0,
                   sh_coeffs: int = 0, extent: float = 4.0):
    """Gaussian cloud of BASELINE.md section 2.2 (all tensors CPU fp32)."""
    g = torch.Generator().manual_seed(seed)
    fx = W / (2.0 * math.tan(0.5 * fovx))
    means3D = (torch.rand(P, 3, generator=g) * 2.0 - 1.0) * extent
    s_px = torch.exp(math.log(sigma_px) + 0.6 * torch.randn(P, 1, generator=g))
    aniso = torch.rand(P, 3, generator=g) + 0.5
    scales = s_px * aniso * 8.0 / fx
    rot = torch.randn(P, 4, generator=g)
    rotations = rot / rot.norm(dim=1, keepdim=True)
    opacities = torch.rand(P, 1, generator=g) * 0.9 + 0.05
    feats = torch.randn(P, K, generator=g)
    colors = feats / feats.norm(dim=1, keepdim=True)
    out = SimpleNamespace(means3D=means3D.contiguous(), scales=scales.contiguous(), rotations=rotations.contiguous(),
                          opacities=opacities.contiguous(), colors=colors.contiguous(), shs=None)
    if sh_coeffs > 0:
        out.shs = (0.3 * torch.randn(P, sh_coeffs, 3, generator=g)).contiguous()
    return out


def make_upstream(K: int, H: int, W: int, seed: int = 1) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(K, H, W, generator=g) / float(H * W)).contiguous()


def scene(P: int, H: int, W: int, K: int, cam: int = 0, seed: int = 0, sh_coeffs: int = 0, sigma_px: float = 2.0):
    """One synthetic scene: camera + Gaussians + upstream gradient (+ mask gradient for the depth variant)."""
    c = make_camera(H, W, cam)
    g = make_gaussians(P, K, W, seed=seed, sh_coeffs=sh_coeffs, sigma_px=sigma_px)
    return SimpleNamespace(cam=c, gauss=g, dL_dout=make_upstream(K, H, W, seed=1),
                           dL_dmask=make_upstream(1, H, W, seed=2), bg=torch.zeros(max(K, 3)), P=P, H=H, W=W, K=K)
