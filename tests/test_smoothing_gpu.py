"""GPU: the fused feature-smoothing op (csrc/smooth.cu) against the reference's own tensor
expression in plain PyTorch fp32 (scene/gaussian_model_ff.py:353-362 + gaussian_renderer/__init__.py:362-363), values
and gradients; tolerance 1e-5 relative (the op is a handful of fp32 operations per element)."""
import sys
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(P, C, Ks, seed, zero_rows=False):
    g = torch.Generator().manual_seed(seed)
    F = torch.randn(P, C, generator=g) * torch.rand(P, 1, generator=g) * 3
    if zero_rows:
        F[::7] = 0.0                              # F.normalize clamps the norm at 1e-12: rows of zeros stay zeros
    idx = torch.randint(0, P, (P, Ks), generator=g)
    idx[:, 0] = torch.arange(P)                   # a KNN map contains the point itself
    idx[5] = idx[5, 0]                            # repeated neighbours in one row
    w = torch.randn(P, C, generator=g)
    return F.cuda(), idx.cuda(), w.cuda()


@pytest.mark.parametrize("P,C,Ks", [(5000, 32, 8), (3000, 16, 4), (2000, 64, 8), (1500, 3, 2), (1000, 32, 16), (64, 32, 8)])
@pytest.mark.parametrize("normalize_output", [False, True])
def test_fused_smoothing_matches_torch_expression(P, C, Ks, normalize_output):
    from seganygaussians_b200.smoothing import smooth_point_features, reference_expression
    F, idx, w = _case(P, C, Ks, seed=P + C)
    Fa, Fb = F.clone().requires_grad_(True), F.clone().requires_grad_(True)
    ya = smooth_point_features(Fa, idx, normalize_output)
    yb = reference_expression(Fb, idx, normalize_output)
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6), float((ya - yb).abs().max())
    (ya * w).sum().backward()
    (yb * w).sum().backward()
    scale = float(Fb.grad.abs().max())
    assert torch.allclose(Fa.grad, Fb.grad, rtol=1e-4, atol=1e-5 * scale), float((Fa.grad - Fb.grad).abs().max())


def test_zero_rows_follow_f_normalize():
    from seganygaussians_b200.smoothing import smooth_point_features, reference_expression
    F, idx, w = _case(2000, 32, 8, seed=9, zero_rows=True)
    Fa, Fb = F.clone().requires_grad_(True), F.clone().requires_grad_(True)
    ya, yb = smooth_point_features(Fa, idx, False), reference_expression(Fb, idx, False)
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6)
    (ya * w).sum().backward(); (yb * w).sum().backward()
    nz = (F.abs().sum(1) > 0)
    assert torch.allclose(Fa.grad[nz], Fb.grad[nz], rtol=1e-4, atol=1e-5 * float(Fb.grad[nz].abs().max()))


def test_drop_in_renderer_uses_the_same_neighbours_as_the_reference_method():
    """`render_contrastive_feature(smooth_type='traditional')` of the drop-in: fused op == the model's own method for the
    same RNG state (both draw `torch.randperm(K)` once)."""
    sys.path.insert(0, os.path.join(ROOT, "seganygaussians_b200", "dropin"))
    import importlib
    gr = importlib.import_module("gaussian_renderer")
    from seganygaussians_b200.smoothing import reference_expression
    P, K = 4000, 16
    g = torch.Generator().manual_seed(1)
    pc = SimpleNamespace(_point_features=torch.randn(P, 32, generator=g).cuda().requires_grad_(True),
                         feature_smooth_map={"K": K, "m": torch.randint(0, P, (P, K), generator=g).cuda()},
                         get_xyz=None)
    torch.manual_seed(123)
    fused = gr._fused_traditional_smoothing(pc, K, 0.5, True)
    torch.manual_seed(123)
    sel = torch.randperm(K)[: int(K * 0.5)]
    ref = reference_expression(pc._point_features, pc.feature_smooth_map["m"][:, sel], True)
    assert torch.allclose(fused, ref, rtol=1e-5, atol=1e-6)


def test_training_step_front_end_end_to_end():
    """What one SAGA training iteration asks of the drop-in (`train_contrastive_feature.py:228`):
    `render_contrastive_feature(cam, feature_gaussians, pipe, bg, norm_point_features=True, smooth_type='traditional',
    smooth_K=16)` on a model without a neighbour map yet -> KNN map through the pytorch3d stand-in, fused smoothing,
    K=32 rasterization, gradients back to `_point_features`.  Checked against the same pipeline assembled by hand from
    the reference's tensor expression and a brute-force neighbour map."""
    import importlib
    import math
    import numpy as np
    import seganygaussians_b200 as S
    from seganygaussians_b200 import synthetic
    from seganygaussians_b200.smoothing import reference_expression
    from oracle import knn_oracle
    S.activate()
    gr = importlib.import_module("gaussian_renderer")
    assert "seganygaussians_b200" in gr.__file__
    P, H, W, K = 6000, 96, 128, 32
    sc = synthetic.scene(P, H, W, K)
    g, c = sc.gauss, sc.cam
    dev = torch.device("cuda")
    feats = (torch.randn(P, K, generator=torch.Generator().manual_seed(4))).to(dev).requires_grad_(True)
    pc = SimpleNamespace(get_xyz=g.means3D.to(dev), get_opacity=g.opacities.to(dev), get_scaling=g.scales.to(dev),
                         get_rotation=g.rotations.to(dev), _point_features=feats, get_point_features=feats,
                         feature_smooth_map=None, active_sh_degree=0)
    cam = SimpleNamespace(FoVx=2 * math.atan(c.tanfovx), FoVy=2 * math.atan(c.tanfovy), feature_height=H, feature_width=W,
                          world_view_transform=c.world_view_transform.to(dev), full_proj_transform=c.full_proj_transform.to(dev),
                          camera_center=c.camera_center.to(dev))
    pipe = SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    bg = torch.zeros(K, device=dev)
    dL = sc.dL_dout.to(dev)

    torch.manual_seed(77)
    out = gr.render_contrastive_feature(cam, pc, pipe, bg, norm_point_features=True, smooth_type="traditional", smooth_K=16)
    (out["render"] * dL).sum().backward()
    got_img, got_grad = out["render"].detach().clone(), feats.grad.detach().clone()
    assert pc.feature_smooth_map["K"] == 16 and pc.feature_smooth_map["m"].shape == (P, 16)
    # the neighbour map is the exact 16-NN map (self first)
    oi, od = knn_oracle.knn_bruteforce(g.means3D.numpy(), None, K=16)
    m = pc.feature_smooth_map["m"].cpu().numpy()
    d_m = ((g.means3D.numpy()[m] - g.means3D.numpy()[:, None, :]) ** 2).sum(-1)
    assert np.allclose(d_m, od, rtol=1e-5, atol=1e-10) and np.array_equal(m[:, 0], np.arange(P))

    feats.grad = None
    torch.manual_seed(77)
    sel = torch.randperm(16)[:8]
    colors = reference_expression(feats, pc.feature_smooth_map["m"][:, sel], True)
    rast = gr.GaussianRasterizerContrastiveF(gr._settings(cam, pc, pipe, bg, 1.0, H, W))
    img, _ = rast(means3D=pc.get_xyz, means2D=torch.zeros_like(pc.get_xyz, requires_grad=True), shs=None, colors_precomp=colors,
                  opacities=pc.get_opacity, scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)
    (img * dL).sum().backward()
    img = img.detach()
    assert torch.allclose(got_img, img, rtol=1e-4, atol=2e-5 * float(img.abs().max()))
    assert torch.allclose(got_grad, feats.grad, rtol=1e-3, atol=2e-5 * float(feats.grad.abs().max()))
