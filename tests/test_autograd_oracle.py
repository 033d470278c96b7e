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


def _direct(sc, K, *, cov=None, scale_modifier=1.0):
    """C oracle and autograd oracle called directly (paths tests.common does not expose: cov3D_precomp, scale_modifier)."""
    from oracle import oracle
    g, c = sc.gauss, sc.cam
    F64 = torch.float64
    leaf = lambda t: t.detach().clone().to(F64).requires_grad_(True)
    means3D, opac, colors = leaf(g.means3D), leaf(g.opacities), leaf(g.colors)
    scales = None if cov is not None else leaf(g.scales)
    rots = None if cov is not None else leaf(g.rotations)
    cov_l = None if cov is None else leaf(cov)
    fw = ag.forward(means3D=means3D, opacities=opac, bg=torch.zeros(3), viewmatrix=c.world_view_transform,
                    projmatrix=c.full_proj_transform, campos=c.camera_center, image_height=sc.H, image_width=sc.W,
                    tanfovx=c.tanfovx, tanfovy=c.tanfovy, scale_modifier=scale_modifier, colors_precomp=colors, scales=scales,
                    rotations=rots, cov3D_precomp=cov_l)
    (fw.color * sc.dL_dout[:K].to(F64)).sum().backward()
    ofw = oracle.forward(means3D=g.means3D.numpy(), opacities=g.opacities.numpy(), bg=np.zeros(3, np.float32),
                         viewmatrix=c.world_view_transform.numpy(), projmatrix=c.full_proj_transform.numpy(),
                         campos=c.camera_center.numpy(), image_height=sc.H, image_width=sc.W, tanfovx=c.tanfovx, tanfovy=c.tanfovy,
                         scale_modifier=scale_modifier, colors_precomp=g.colors.numpy(),
                         scales=None if cov is not None else g.scales.numpy(), rotations=None if cov is not None else g.rotations.numpy(),
                         cov3D_precomp=None if cov is None else cov.numpy(), num_channels=K)
    obw = oracle.backward(ofw, sc.dL_dout[:K].numpy())
    return fw, (means3D, opac, colors, scales, rots, cov_l), ofw, obw


def _close(a, b, what):
    r, d, s = common.float_err(a, b)
    assert r <= 1.0, f"{what}: max|d|={d:.3e} max|ref|={s:.3e} tol-ratio={r:.3f}"


def test_cov3d_precomp_gradient():
    """dL_dcov3D (only observable on the cov3D_precomp path; off-diagonals carry the factor 2, A.19)."""
    P, H, W, K = 300, 48, 64, 3
    sc = synthetic.scene(P, H, W, K, sigma_px=3.0)
    g = sc.gauss
    Sg = ag._cov3d(g.scales.to(torch.float64), g.rotations.to(torch.float64), 1.0)
    cov = torch.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]], dim=1).to(torch.float32).contiguous()
    fw, (means3D, opac, colors, _, _, cov_l), ofw, obw = _direct(sc, K, cov=cov)
    assert ofw.num_rendered == fw.num_rendered and np.array_equal(ofw.point_list, fw.point_list)
    _close(fw.color.detach().numpy(), ofw.color, "color")
    _close(cov_l.grad.numpy(), obw.cov3D, "dL_dcov3D")
    _close(means3D.grad.numpy(), obw.means3D, "dL_dmeans3D")
    _close(opac.grad.numpy(), obw.opacity, "dL_dopacity")
    assert np.abs(obw.cov3D).max() > 0 and np.abs(obw.scales).max() == 0      # no scale / rotation gradient on this path


def test_scale_modifier():
    """scale_modifier != 1: the reference's dL_dscales lacks the factor scale_modifier (CF backward.cu:297-325; the viewer's
    scaling slider is the only caller with a value other than 1) -- the C oracle and the CUDA path keep that, and the autograd
    oracle models it as a stop-gradient; dL_drotations and dL_dmeans3D are true derivatives."""
    P, H, W, K = 300, 48, 64, 3
    sc = synthetic.scene(P, H, W, K, sigma_px=2.0)
    fw, (means3D, opac, colors, scales, rots, _), ofw, obw = _direct(sc, K, scale_modifier=1.7)
    assert ofw.num_rendered == fw.num_rendered and np.array_equal(ofw.radii, fw.radii)
    _close(fw.color.detach().numpy(), ofw.color, "color")
    _close(scales.grad.numpy(), obw.scales, "dL_dscales")
    _close(rots.grad.numpy(), obw.rotations, "dL_drotations")
    _close(means3D.grad.numpy(), obw.means3D, "dL_dmeans3D")


def test_random_sweep_cameras_shapes_and_conventions():
    """24 seeded random configurations: camera position (also INSIDE the cloud: near-plane culls, splats thousands of pixels
    wide), field of view, ragged image sizes, K in {3, 5, 32}, SH degrees 0-3, DEPTH variant, background, unnormalised
    quaternions.  Integer state exact; floats within 5x the suite tolerance (the C oracle is fp32 like the reference, and a
    splat at view depth 0.37 with a 1269-pixel radius loses four digits to cancellation in its mean gradient)."""
    rng = np.random.RandomState(123)
    for t in range(24):
        P, H, W = int(rng.randint(50, 400)), int(rng.randint(17, 70)), int(rng.randint(17, 90))
        K, cam, fovx = int(rng.choice([3, 3, 5, 32])), int(rng.randint(0, 8)), float(rng.uniform(0.5, 1.7))
        use_sh = (K == 3) and rng.rand() < 0.5
        deg = int(rng.randint(0, 4)) if use_sh else 0
        depth = (K == 3) and rng.rand() < 0.4
        sigma = float(rng.choice([1.0, 2.0, 5.0, 15.0]))
        sc = synthetic.scene(P, H, W, K, cam=cam, sh_coeffs=16 if use_sh else 0, sigma_px=sigma, seed=t)
        sc.cam = synthetic.make_camera(H, W, cam, fovx=fovx, radius=float(rng.uniform(2.0, 9.0)))
        sc.gauss.rotations = (sc.gauss.rotations * torch.tensor(rng.uniform(0.6, 1.4, (P, 1)), dtype=torch.float32)).contiguous()
        bg = torch.tensor(rng.uniform(0, 1, max(K, 3)).astype(np.float32)) if rng.rand() < 0.5 else None
        a = ag.run_scene(sc, K, depth=depth, use_sh=use_sh, sh_degree=deg, bg=bg)
        o = common.run_oracle(sc, K, depth=depth, use_sh=use_sh, sh_degree=deg, bg=bg)
        tag = f"case {t}: P={P} {H}x{W} K={K} cam={cam} fovx={fovx:.2f} sh={use_sh}/{deg} depth={depth} sigma={sigma}"
        for name in common.INT_FWD:
            x, y = getattr(a, name, None), getattr(o, name, None)
            if x is not None and y is not None:
                assert np.array_equal(np.asarray(x), np.asarray(y)), (tag, name)
        for name in common.FLOAT_FWD + common.GRADS:
            x, y = getattr(a, name, None), getattr(o, name, None)
            if x is not None and y is not None:
                r, d, s = common.float_err(x, y)
                assert r <= 5.0, f"{tag}: {name} max|d|={d:.3e} max|ref|={s:.3e} tol-ratio={r:.2f}"
