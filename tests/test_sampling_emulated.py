"""The fused loss-side consumer (csrc/sample_kernels.cuh: pixel-norm sum + bilinear sampling of the rays, and its backward) executed
on the CPU under the CUDA execution shim against the reference's tensor expression (train_contrastive_feature.py:234-254) and its
autograd gradient.  The same kernels run on the GPU in tests/test_sampling_gpu.py."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from seganygaussians_b200 import sampling

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    d = tempfile.mkdtemp(prefix="sagars_emu_")
    so = os.path.join(d, "libemu_sample.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "tests", "cuda_emu"), "-I", os.path.join(ROOT, "seganygaussians_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cuda_emu", "emu_sample.cpp"), "-o", so])
    L = C.CDLL(so)
    L.emu_sample_forward.restype = C.c_int
    L.emu_sample_forward.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.emu_sample_backward.restype = C.c_int
    L.emu_sample_backward.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 3
    return L


@pytest.mark.parametrize("case", [("down", 8, 23, 37, 11, 19, 60), ("up", 5, 9, 13, 20, 31, 80), ("same", 32, 16, 24, 16, 24, 50),
                                  ("one_ray", 3, 7, 7, 14, 5, 1), ("no_rays", 4, 6, 6, 6, 6, 0)], ids=lambda c: c[0])
def test_emulated_sample_rays_matches_the_tensor_expression(emu, case):
    name, Cn, H, W, h, w, S = case
    g = torch.Generator().manual_seed(3)
    img = torch.randn(Cn, H, W, generator=g)
    img[:, 0, 0] = 0.0                                   # a zero-norm pixel: its norm gradient is 0, not NaN
    rays = torch.randperm(h * w, generator=g)[:S].sort().values
    img_t = img.clone().requires_grad_(True)
    want_s, want_n = sampling.reference_expression(img_t, (h, w), rays)
    g_s = torch.randn(Cn, S, generator=g)
    g_n = torch.tensor(0.7)
    ((want_s * g_s).sum() + want_n * g_n).backward()

    p = lambda a: a.ctypes.data
    im = np.ascontiguousarray(img.numpy())
    ry = np.ascontiguousarray(rays.numpy().astype(np.int64))
    out = np.zeros((Cn, max(S, 1)), np.float32)
    ns = np.zeros(1, np.float32)
    assert emu.emu_sample_forward(Cn, H, W, h, w, p(im), p(ry), S, p(out), p(ns)) == 0
    np.testing.assert_allclose(ns[0] / (H * W), float(want_n.detach()), rtol=1e-5)
    if S:
        np.testing.assert_allclose(out[:, :S], want_s.detach().numpy(), rtol=1e-5, atol=2e-5)   # four-tap sums of O(1) values: order of the fp32 lerp
    gi = np.full((Cn, H, W), np.nan, np.float32)
    gs = np.ascontiguousarray(g_s.numpy()) if S else np.zeros((Cn, 1), np.float32)
    gn = np.array([0.7], np.float32)
    assert emu.emu_sample_backward(Cn, H, W, h, w, p(im), p(ry), S, p(gs), p(gn), p(gi)) == 0
    np.testing.assert_allclose(gi, img_t.grad.numpy(), rtol=2e-5, atol=2e-5)
