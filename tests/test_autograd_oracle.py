"""The C oracle's hand-restated backward against gradients DERIVED by torch.autograd from an independent float64
restatement of the forward only (oracle/autograd_oracle.py).  CPU only.

The golden vectors pin the C oracle to the reference's outputs on a handful of scenes; this test pins its analytic
gradient chain (blend -> conic -> cov2D -> cov3D -> scale / rotation, mean2D -> mean3D, SH, mask) to calculus on
scenes chosen to hit the special cases: field-of-view-clamped Gaussians that reach the screen, non-zero background
with translucent pixels, SH colours with clamped channels, the DEPTH variant's mask / depth outputs.
Tolerance: the suite-wide 1e-4 relative (tests/common.py); integer state exact."""
import numpy as np
import pytest
import torch

from tests import common
from seganygaussians_b200 import synthetic
from oracle import autograd_oracle as ag

CASES = [
    # id,            P,   H,  W,  K, depth, sh, deg, sigma_px, bg
    ("cf_k32",       400, 48, 64, 32, False, False, 0, 3.0, None),
    ("base_sh3",     300, 40, 56, 3, False, True, 3, 3.0, None),
    ("depth_mask",   300, 40, 56, 3, True, False, 0, 3.0, None),
    ("depth_sh2",    500, 64, 48, 3, True, True, 2, 3.0, None),
    ("fovclamp_bg",  200, 48, 64, 3, False, False, 0, 40.0, (0.3, 0.7, 0.1)),
    ("sparse_bg",    60, 33, 47, 3, False, False, 0, 2.0, (0.9, 0.2, 0.5)),     # ragged image size, mostly background
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_c_oracle_gradients_equal_autograd(case):
    _, P, H, W, K, depth, use_sh, deg, sigma, bg = case
    sc = synthetic.scene(P, H, W, K, sh_coeffs=16 if use_sh else 0, sigma_px=sigma)
    bg_t = None if bg is None else torch.tensor(bg)
    a = ag.run_scene(sc, K, depth=depth, use_sh=use_sh, sh_degree=deg, bg=bg_t)
    o = common.run_oracle(sc, K, depth=depth, use_sh=use_sh, sh_degree=deg, bg=bg_t)
    assert o.num_rendered > 0
    ok, lines = common.compare(a, o, verbose=False)
    assert ok, "\n".join(lines)
    # the comparison must have covered every gradient the variant has
    names = [l.split()[1] for l in lines if l.strip().startswith("FLOAT")]
    for need in ("g_means3D", "g_means2D", "g_opacity", "g_scales", "g_rotations", "g_sh" if use_sh else "g_colors"):
        assert need in names, (need, names)
    if depth:
        assert "g_mask" in names and "out_depth" in names


def test_fov_clamped_gaussians_are_on_screen_in_the_clamp_case():
    """Guard for the case above: it only tests A.20 if clamped Gaussians actually contribute."""
    P, H, W, K = 200, 48, 64, 3
    sc = synthetic.scene(P, H, W, K, sigma_px=40.0)
    o = common.run_oracle(sc, K, backward=False)
    t = (torch.cat([sc.gauss.means3D, torch.ones(P, 1)], 1) @ sc.cam.world_view_transform)[:, :3]
    clamped = ((t[:, 0] / t[:, 2]).abs() > 1.3 * sc.cam.tanfovx) | ((t[:, 1] / t[:, 2]).abs() > 1.3 * sc.cam.tanfovy)
    assert int((clamped.numpy() & (o.radii > 0)).sum()) >= 5


def test_straight_through_clamp_is_exercised():
    """Opaque Gaussians at their centres hit alpha = 0.99: autograd with a plain clamp would give zero there."""
    P, H, W, K = 150, 32, 32, 3
    sc = synthetic.scene(P, H, W, K, sigma_px=4.0)
    sc.gauss.opacities = torch.full_like(sc.gauss.opacities, 0.999)
    a = ag.run_scene(sc, K)
    o = common.run_oracle(sc, K)
    ok, lines = common.compare(a, o, verbose=False)
    assert ok, "\n".join(lines)
    assert np.abs(o.g_opacity).max() > 0
