"""Fused feature smoothing (``sagars_smooth_forward`` / ``sagars_smooth_backward``; kernels in ``csrc/smooth.cu``).

The reference forms the per-Gaussian features of a SAGA training step with five tensor passes and a ``[P, Ks, C]`` gather
(``scene/gaussian_model_ff.py:338-364``, ``gaussian_renderer/__init__.py:362-363``); this is the same function as one
autograd op in front of the rasterizer.  CUDA tensors only, no fallback.
"""
from __future__ import annotations

import torch

from . import _lib


class _SmoothPointFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, point_features, nbr_idx, normalize_output):
        lib = _lib.load()
        if point_features.dim() != 2 or nbr_idx.dim() != 2 or nbr_idx.shape[0] != point_features.shape[0]:
            raise RuntimeError("point_features must be [P, C] and nbr_idx [P, Ks]")
        if not point_features.is_cuda:
            raise RuntimeError("point_features must be a CUDA tensor (libsagars has no CPU path)")
        dev = point_features.device
        F = point_features.detach().to(torch.float32).contiguous()
        idx = nbr_idx.to(device=dev, dtype=torch.int64).contiguous()
        P, C, Ks = int(F.shape[0]), int(F.shape[1]), int(idx.shape[1])
        norm = bool(normalize_output)
        with torch.cuda.device(dev):
            out = torch.empty_like(F)
            mean_norm = torch.empty((P,), dtype=torch.float32, device=dev) if norm else None
            rc = lib.sagars_smooth_forward(dev.index if dev.index is not None else torch.cuda.current_device(), P, C, Ks,
                                           F.data_ptr(), idx.data_ptr(), 1 if norm else 0, out.data_ptr(),
                                           None if mean_norm is None else mean_norm.data_ptr(),
                                           int(torch.cuda.current_stream(dev).cuda_stream))
            _lib.check(rc)
        ctx.normalize_output = norm
        ctx.save_for_backward(F, idx, mean_norm if norm else torch.empty(0, device=dev), out if norm else torch.empty(0, device=dev))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        F, idx, mean_norm, out = ctx.saved_tensors
        dev = F.device
        P, C, Ks = int(F.shape[0]), int(F.shape[1]), int(idx.shape[1])
        g = grad_out.to(torch.float32).contiguous()
        norm = ctx.normalize_output
        with torch.cuda.device(dev):
            scratch = torch.empty_like(F)
            dF = torch.empty_like(F)
            rc = lib.sagars_smooth_backward(dev.index if dev.index is not None else torch.cuda.current_device(), P, C, Ks,
                                            F.data_ptr(), idx.data_ptr(), 1 if norm else 0,
                                            mean_norm.data_ptr() if norm else None, out.data_ptr() if norm else None,
                                            g.data_ptr(), scratch.data_ptr(), dF.data_ptr(),
                                            int(torch.cuda.current_stream(dev).cuda_stream))
            _lib.check(rc)
        return dF, None, None


def smooth_point_features(point_features: torch.Tensor, nbr_idx: torch.Tensor, normalize_output: bool = False) -> torch.Tensor:
    """``normalize(point_features)[nbr_idx].mean(1)`` (+ ``x / (||x|| + 1e-9)`` when ``normalize_output``), differentiable
    with respect to ``point_features``.  ``nbr_idx``: ``[P, Ks]`` integer tensor of row indices."""
    return _SmoothPointFeatures.apply(point_features, nbr_idx, normalize_output)


def reference_expression(point_features: torch.Tensor, nbr_idx: torch.Tensor, normalize_output: bool = False) -> torch.Tensor:
    """The reference's own tensor expression (gaussian_model_ff.py:353-362, gaussian_renderer/__init__.py:362-363): the
    plain-PyTorch fp32 reference the fused op is tested against."""
    normed = torch.nn.functional.normalize(point_features, dim=-1, p=2)
    ret = normed[nbr_idx, :].mean(dim=1)
    if normalize_output:
        ret = ret / (ret.norm(dim=1, keepdim=True) + 1e-9)
    return ret
