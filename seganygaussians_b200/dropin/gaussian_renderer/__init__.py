"""Drop-in for the reference's render binding ``gaussian_renderer/__init__.py`` (lines 18-382): the four
entry points ``render``, ``render_mask``, ``render_with_depth`` and ``render_contrastive_feature`` with the
same signatures, the same duck-typed camera / model / pipe attributes and the same result dictionaries, but
bound to the libsagars-backed rasterizers.

Unlike the reference module this one imports nothing from ``scene`` / ``utils`` (they were only used for
type annotations and for the optional Python SH evaluation, restated here), so it can be imported without
plyfile / simple_knn / pytorch3d.
"""
import math

import torch

from seganygaussians_b200.rasterizer import (
    GaussianRasterizationSettings,
    GaussianRasterizer,
    GaussianRasterizerContrastiveF,
    GaussianRasterizerDepth,
)

_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)


def eval_sh(deg, sh, dirs):
    """Real SH basis (degree <= 3) contracted with coefficients ``sh[..., C, (deg+1)**2]`` at unit ``dirs[..., 3]``
    (behaviour of the reference's ``utils/sh_utils.py:57-112`` for deg 0..3)."""
    assert 0 <= deg <= 3 and sh.shape[-1] >= (deg + 1) ** 2
    x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
    basis = [torch.full_like(x, _SH_C0)]
    if deg > 0:
        basis += [-_SH_C1 * y, _SH_C1 * z, -_SH_C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        basis += [_SH_C2[0] * xy, _SH_C2[1] * yz, _SH_C2[2] * (2.0 * zz - xx - yy), _SH_C2[3] * xz, _SH_C2[4] * (xx - yy)]
    if deg > 2:
        basis += [_SH_C3[0] * y * (3 * xx - yy), _SH_C3[1] * xy * z, _SH_C3[2] * y * (4 * zz - xx - yy),
                  _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _SH_C3[4] * x * (4 * zz - xx - yy),
                  _SH_C3[5] * z * (xx - yy), _SH_C3[6] * x * (xx - 3 * yy)]
    out = 0
    for k, b in enumerate(basis):
        out = out + b * sh[..., k]
    return out


def _screenspace_points(pc):
    # dummy tensor whose .grad receives the screen-space mean gradients (reference :26-30)
    pts = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda") + 0
    try:
        pts.retain_grad()
    except Exception:
        pass
    return pts


def _settings(cam, pc, pipe, bg_color, scaling_modifier, height, width):
    return GaussianRasterizationSettings(
        image_height=int(height), image_width=int(width),
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5),
        bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=pc.active_sh_degree, campos=cam.camera_center,
        prefiltered=False, debug=pipe.debug)


def _covariance_inputs(pc, pipe, scaling_modifier):
    if pipe.compute_cov3D_python:
        return None, None, pc.get_covariance(scaling_modifier)
    return pc.get_scaling, pc.get_rotation, None


def _colour_inputs(cam, pc, pipe, override_color):
    """(shs, colors_precomp): SH evaluated by the rasterizer, or in Python when the pipe asks for it."""
    if override_color is not None:
        return None, override_color
    if pipe.convert_SHs_python:
        feats = pc.get_features
        shs_view = feats.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
        dir_pp = pc.get_xyz - cam.camera_center.repeat(feats.shape[0], 1)
        dir_pp = dir_pp / dir_pp.norm(dim=1, keepdim=True)
        return None, torch.clamp_min(eval_sh(pc.active_sh_degree, shs_view, dir_pp) + 0.5, 0.0)
    return pc.get_features, None


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None, filtered_mask=None):
    """RGB render (reference ``render``, :18-104). ``bg_color`` must be on the GPU."""
    screenspace_points = _screenspace_points(pc)
    rasterizer = GaussianRasterizer(_settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier,
                                              viewpoint_camera.image_height, viewpoint_camera.image_width))
    opacity = pc.get_opacity
    if filtered_mask is not None:
        opacity = opacity.detach().clone()
        opacity[filtered_mask, :] = 0
    scales, rotations, cov3D_precomp = _covariance_inputs(pc, pipe, scaling_modifier)
    shs, colors_precomp = _colour_inputs(viewpoint_camera, pc, pipe, override_color)
    rendered_image, radii = rasterizer(means3D=pc.get_xyz, means2D=screenspace_points, shs=shs,
                                       colors_precomp=colors_precomp, opacities=opacity, scales=scales,
                                       rotations=rotations, cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}


def render_mask(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, precomputed_mask=None):
    """Per-Gaussian mask rendered as a 3-channel colour (reference ``render_mask``, :108-190)."""
    screenspace_points = _screenspace_points(pc)
    rasterizer = GaussianRasterizer(_settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier,
                                              viewpoint_camera.image_height, viewpoint_camera.image_width))
    mask = pc.get_mask if precomputed_mask is None else precomputed_mask
    if len(mask.shape) == 1 or mask.shape[-1] == 1:
        mask = mask.squeeze().unsqueeze(-1).repeat([1, 3]).cuda()
    scales, rotations, cov3D_precomp = _covariance_inputs(pc, pipe, scaling_modifier)
    rendered_mask, radii = rasterizer(means3D=pc.get_xyz, means2D=screenspace_points, shs=None, colors_precomp=mask,
                                      opacities=pc.get_opacity, scales=scales, rotations=rotations,
                                      cov3D_precomp=cov3D_precomp)
    return {"mask": rendered_mask, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}


def render_with_depth(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None,
                      override_mask=None, filtered_mask=None):
    """RGB + mask + view depth (reference ``render_with_depth``, :194-294)."""
    screenspace_points = _screenspace_points(pc)
    rasterizer = GaussianRasterizerDepth(_settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier,
                                                   viewpoint_camera.image_height, viewpoint_camera.image_width))
    opacity = pc.get_opacity
    if filtered_mask is not None:
        opacity = opacity.detach().clone()
        opacity[filtered_mask, :] = -1.
    mask = pc.get_mask if override_mask is None else override_mask
    scales, rotations, cov3D_precomp = _covariance_inputs(pc, pipe, scaling_modifier)
    shs, colors_precomp = _colour_inputs(viewpoint_camera, pc, pipe, override_color)
    rendered_image, rendered_mask, rendered_depth, radii = rasterizer(
        means3D=pc.get_xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp, opacities=opacity,
        mask=mask, scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image, "mask": rendered_mask, "depth": rendered_depth,
            "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii}


def _fused_traditional_smoothing(pc, K, dropout, normalize_output):
    """``pc.get_smoothed_point_features(K, dropout)`` (+ the renderer's renormalisation) as ONE fused op
    (``seganygaussians_b200.smoothing``).  Same neighbour map, same ``torch.randperm`` draw as the reference
    (scene/gaussian_model_ff.py:338-364), so the selected neighbours are identical.  Returns None when the model does
    not look like the reference's FeatureGaussianModel (the caller then uses the model's own method)."""
    if K <= 1 or not hasattr(pc, "feature_smooth_map") or not hasattr(pc, "_point_features"):
        return None
    if not (0 < dropout < 1) or int(K * dropout) < 1 or not pc._point_features.is_cuda:
        return None
    from seganygaussians_b200.smoothing import smooth_point_features
    with torch.no_grad():
        if pc.feature_smooth_map is None or pc.feature_smooth_map["K"] != K:
            import pytorch3d.ops
            xyz = pc.get_xyz
            nearest_k_idx = pytorch3d.ops.knn_points(xyz.unsqueeze(0), xyz.unsqueeze(0), K=K).idx.squeeze()
            pc.feature_smooth_map = {"K": K, "m": nearest_k_idx}
    select_point = torch.randperm(K)[: int(K * dropout)]
    select_idx = pc.feature_smooth_map["m"][:, select_point]
    return smooth_point_features(pc._point_features, select_idx, normalize_output)


def render_contrastive_feature(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, norm_point_features=False,
                               smooth_type=None, smooth_weights=None, smooth_K=16):
    """K-dim affinity-feature render at the camera's feature resolution (reference
    ``render_contrastive_feature``, :300-382)."""
    screenspace_points = _screenspace_points(pc)
    rasterizer = GaussianRasterizerContrastiveF(_settings(viewpoint_camera, pc, pipe, bg_color, scaling_modifier,
                                                          viewpoint_camera.feature_height,
                                                          viewpoint_camera.feature_width))
    scales, rotations, cov3D_precomp = _covariance_inputs(pc, pipe, scaling_modifier)
    colors_precomp = None
    if smooth_type is None:
        colors_precomp = pc.get_point_features
    elif smooth_type == 'multi_res':
        colors_precomp = pc.get_multi_resolution_smoothed_point_features(smooth_weights=smooth_weights)
    elif smooth_type == 'traditional':
        fused = _fused_traditional_smoothing(pc, smooth_K, 0.5, norm_point_features)
        if fused is not None:
            colors_precomp, norm_point_features = fused, False      # the fused op already renormalised
        else:
            colors_precomp = pc.get_smoothed_point_features(K=smooth_K, dropout=0.5)
    if norm_point_features:
        colors_precomp = colors_precomp / (colors_precomp.norm(dim=1, keepdim=True) + 1e-9)
    rendered_image, radii = rasterizer(means3D=pc.get_xyz, means2D=screenspace_points, shs=None,
                                       colors_precomp=colors_precomp, opacities=pc.get_opacity, scales=scales,
                                       rotations=rotations, cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}
