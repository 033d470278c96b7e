"""GPU parity of the fused loss-side consumer (seganygaussians_b200.sampling.sample_rays) against the
reference's tensor expression (train_contrastive_feature.py:232-254) in plain PyTorch fp32: values 1e-5, gradients 1e-4 relative."""
import pytest
import torch

from seganygaussians_b200 import sampling

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [("training_like", 32, 270, 480, 256, 455, 1000, "mask"), ("upsample", 32, 65, 100, 270, 480, 777, "index"),
                                  ("k3", 3, 128, 96, 128, 96, 500, "mask"), ("no_rays", 8, 32, 48, 16, 24, 0, "index")], ids=lambda c: c[0])
def test_sample_rays_matches_the_reference_expression(case):
    name, C, H, W, h, w, S, kind = case
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    img = torch.randn(C, H, W, generator=g).to(dev)
    img[:, 3, 5] = 0.0
    flat = torch.randperm(h * w, generator=g)[:S].sort().values.to(dev)
    if kind == "mask":
        rays = torch.zeros(h * w, dtype=torch.bool, device=dev)
        rays[flat] = True
        rays = rays.reshape(h, w)
    else:
        rays = flat
    a = img.clone().requires_grad_(True)
    b = img.clone().requires_grad_(True)
    want_s, want_n = sampling.reference_expression(a, (h, w), rays)
    got_s, got_n = sampling.sample_rays(b, (h, w), rays)
    assert got_s.shape == want_s.shape
    torch.testing.assert_close(got_s, want_s, rtol=1e-5, atol=2e-5)      # four-tap sums of O(1) values: order of the fp32 lerp
    torch.testing.assert_close(got_n, want_n, rtol=1e-5, atol=0)
    gs = torch.randn(C, S, generator=g).to(dev)
    ((want_s * gs).sum() + (1 - want_n) ** 2).backward()
    ((got_s * gs).sum() + (1 - got_n) ** 2).backward()
    torch.testing.assert_close(b.grad, a.grad, rtol=1e-4, atol=2e-6)


def test_sample_rays_rejects_cpu_tensors():
    with pytest.raises(RuntimeError):
        sampling.sample_rays(torch.zeros(3, 4, 4), (4, 4), torch.zeros(1, dtype=torch.long))
