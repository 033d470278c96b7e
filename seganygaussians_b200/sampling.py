"""Fused loss-side consumer of the K-feature render (``sagars_sample_rays_forward`` / ``_backward``; kernels in
``csrc/sample.cu``) -- SURVEY.md section 8(f) rank 3.

The reference's training step (``train_contrastive_feature.py:232-254``) takes the ``[32, H, W]`` render, computes the mean pixel
norm (a regulariser), resizes the whole image bilinearly to the mask resolution and then reads ~1000 sampled rays from it.  This
module is the same function as one autograd op that never materialises the resized image.  CUDA tensors only, no fallback.
"""
from __future__ import annotations

import torch

from . import _lib


class _SampleRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, out_h, out_w, ray_index):
        lib = _lib.load()
        if render.dim() != 3:
            raise RuntimeError("render must be [C, H, W]")
        if not render.is_cuda:
            raise RuntimeError("render must be a CUDA tensor (libsagars has no CPU path)")
        dev = render.device
        img = render.detach().to(torch.float32).contiguous()
        idx = ray_index.to(device=dev, dtype=torch.int64).contiguous()
        C, H, W = (int(v) for v in img.shape)
        S = int(idx.numel())
        if S and (int(idx.min()) < 0 or int(idx.max()) >= out_h * out_w):
            raise RuntimeError("ray_index out of range for the resized image")
        with torch.cuda.device(dev):
            samples = torch.empty((C, S), dtype=torch.float32, device=dev)
            norm_sum = torch.empty((1,), dtype=torch.float32, device=dev)
            _lib.check(lib.sagars_sample_rays_forward(dev.index if dev.index is not None else torch.cuda.current_device(), C, H, W,
                                                      int(out_h), int(out_w), img.data_ptr(), idx.data_ptr(), S, samples.data_ptr(),
                                                      norm_sum.data_ptr(), int(torch.cuda.current_stream(dev).cuda_stream)))
        ctx.dims = (C, H, W, int(out_h), int(out_w), S)
        ctx.save_for_backward(img, idx)
        return samples, (norm_sum / float(H * W)).reshape(())

    @staticmethod
    def backward(ctx, g_samples, g_norm):
        lib = _lib.load()
        img, idx = ctx.saved_tensors
        C, H, W, out_h, out_w, S = ctx.dims
        dev = img.device
        gs = (torch.zeros((C, S), device=dev) if g_samples is None else g_samples).to(torch.float32).contiguous()
        gn = (torch.zeros((), device=dev) if g_norm is None else g_norm).to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            grad = torch.empty_like(img)
            _lib.check(lib.sagars_sample_rays_backward(dev.index if dev.index is not None else torch.cuda.current_device(), C, H, W,
                                                       out_h, out_w, img.data_ptr(), idx.data_ptr(), S, gs.data_ptr(), gn.data_ptr(),
                                                       grad.data_ptr(), int(torch.cuda.current_stream(dev).cuda_stream)))
        return grad, None, None, None


def sample_rays(render: torch.Tensor, size, sampled_ray: torch.Tensor):
    """``(samples[C, S], mean pixel norm)`` of a ``[C, H, W]`` render:

        norm    = render.norm(dim=0, p=2).mean()
        samples = F.interpolate(render[None], size, mode='bilinear')[0][:, sampled_ray]      # bool mask [h, w] or flat indices

    differentiable with respect to ``render`` (train_contrastive_feature.py:234-254 reads the rays through a boolean mask:
    row-major order of the set positions, which is what ``nonzero`` gives)."""
    h, w = int(size[0]), int(size[1])
    if sampled_ray.dtype == torch.bool:
        if tuple(sampled_ray.shape[-2:]) != (h, w):
            raise RuntimeError("the boolean ray mask must have the resized image's shape")
        idx = torch.nonzero(sampled_ray.reshape(-1), as_tuple=False).reshape(-1)
    else:
        idx = sampled_ray.reshape(-1)
    return _SampleRays.apply(render, h, w, idx)


def reference_expression(render: torch.Tensor, size, sampled_ray: torch.Tensor):
    """The reference's own tensor expression (train_contrastive_feature.py:234-237, 250): the plain-PyTorch fp32 reference the fused
    op is tested against."""
    norm = render.norm(dim=0, p=2).mean()
    up = torch.nn.functional.interpolate(render.unsqueeze(0), tuple(int(v) for v in size), mode="bilinear").squeeze(0)
    if sampled_ray.dtype == torch.bool:
        return up[:, sampled_ray], norm
    return up.reshape(up.shape[0], -1)[:, sampled_ray.reshape(-1)], norm
