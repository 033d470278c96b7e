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
