"""the drop-in ``gaussian_renderer`` against the REFERENCE's own ``gaussian_renderer``.

Both bindings are driven with the same duck-typed camera / model / pipe objects: the reference's
``render``, ``render_mask``, ``render_with_depth`` and ``render_contrastive_feature`` run on top of the reference's own CUDA
extensions (``oracle/_ref``, installed by ``oracle/build_ref.py``), ours on top of libsagars; result dictionaries and the
gradients that flow back into the model's leaves are compared (integer outputs exact, fp32 within the parity tolerance).
Also the DEPTH variant's mask-only API (``GaussianRasterizer.forward_mask``) against the reference's own ``forward_mask``."""
import importlib
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests import common
from seganygaussians_b200 import synthetic

pytestmark = pytest.mark.gpu
ROOT = common.ROOT
REF_RENDERER = os.path.join(common.REF_DIR, "renderer")


@pytest.fixture(autouse=True)
def _restore_import_state():
    """The reference's python packages are put in front of sys.path for these tests only: afterwards the path entry and the
    modules imported from it are removed again (other tests import the DROP-IN ``gaussian_renderer`` by that same name)."""
    saved_path = list(sys.path)
    yield
    sys.path[:] = saved_path
    for k in [k for k in list(sys.modules) if k.split(".")[0] in ("gaussian_renderer", "scene", "utils", "arguments")]:
        f = getattr(sys.modules[k], "__file__", None) or ""
        if os.path.realpath(f).startswith(os.path.realpath(common.REF_DIR)) or not f:
            del sys.modules[k]


def _load_bindings():
    """(reference gaussian_renderer bound to the reference extensions, our drop-in bound to libsagars)."""
    if not (os.path.exists(os.path.join(REF_RENDERER, "gaussian_renderer", "__init__.py")) and
            all(common.have_ref(v) for v in ("base", "cf", "depth"))):
        pytest.skip("oracle/_ref (extensions + renderer binding) not built: python oracle/build_ref.py where /root/reference is mounted")
    import seganygaussians_b200 as S
    if S.SHIMS_DIR not in sys.path:
        sys.path.append(S.SHIMS_DIR)                      # plyfile / pytorch3d stand-ins for the reference's `scene` package
    for p in (REF_RENDERER, common.REF_DIR):              # the reference's python packages and its extensions win
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in ("gaussian_renderer", "scene", "utils", "arguments"):
        mod = sys.modules.get(name)
        if mod is not None and not os.path.realpath(getattr(mod, "__file__", None) or "/").startswith(os.path.realpath(common.REF_DIR)):
            for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[k]
    ref = importlib.import_module("gaussian_renderer")
    assert os.path.realpath(ref.__file__).startswith(os.path.realpath(REF_RENDERER)), ref.__file__
    assert os.path.realpath(sys.modules["diff_gaussian_rasterization"].__file__).startswith(os.path.realpath(common.REF_DIR))
    spec = importlib.util.spec_from_file_location("sagars_gaussian_renderer",
                                                  os.path.join(ROOT, "seganygaussians_b200", "dropin", "gaussian_renderer", "__init__.py"))
    ours = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ours)
    return ref, ours


class _Model:
    """What the render bindings read from a Gaussian model (reference scene/gaussian_model.py, gaussian_model_ff.py)."""

    def __init__(self, sc, K, dev, sh_degree=3):
        g = sc.gauss
        leaf = lambda t: t.clone().to(dev).requires_grad_(True)
        self._xyz, self._opacity, self._scaling, self._rotation = leaf(g.means3D), leaf(g.opacities), leaf(g.scales), leaf(g.rotations)
        gen = torch.Generator().manual_seed(5)
        self._sh = leaf(torch.randn(sc.P, (sh_degree + 1) ** 2, 3, generator=gen) * 0.3)
        # [P, 1]: the reference's depth rasterizer returns dL_dmask as [P, 1], which autograd only accepts for a mask of that shape
        self._mask = leaf(torch.rand(sc.P, 1, generator=gen) * 0.5 + 0.5)
        self._point_features = leaf(torch.nn.functional.normalize(torch.randn(sc.P, K, generator=gen), dim=1))
        self.active_sh_degree = self.max_sh_degree = sh_degree

    get_xyz = property(lambda s: s._xyz)
    get_opacity = property(lambda s: s._opacity)
    get_scaling = property(lambda s: s._scaling)
    get_rotation = property(lambda s: s._rotation)
    get_features = property(lambda s: s._sh)
    get_mask = property(lambda s: s._mask)
    get_point_features = property(lambda s: s._point_features)

    def get_covariance(self, scaling_modifier=1):
        # the reference's build_covariance_from_scaling_rotation (scene/gaussian_model.py:33-37): L = R S, Sigma = L L^T, 6 unique
        from utils.general_utils import build_scaling_rotation, strip_symmetric
        L = build_scaling_rotation(scaling_modifier * self._scaling, self._rotation)
        return strip_symmetric(L @ L.transpose(1, 2))

    def leaves(self):
        return {"xyz": self._xyz, "opacity": self._opacity, "scaling": self._scaling, "rotation": self._rotation, "sh": self._sh,
                "mask": self._mask, "features": self._point_features}

    def zero_grad(self):
        for t in self.leaves().values():
            t.grad = None


def _camera(sc, dev):
    c = sc.cam
    import math
    return SimpleNamespace(FoVx=2 * math.atan(c.tanfovx), FoVy=2 * math.atan(c.tanfovy), image_height=sc.H, image_width=sc.W,
                           feature_height=sc.H, feature_width=sc.W, world_view_transform=c.world_view_transform.to(dev),
                           full_proj_transform=c.full_proj_transform.to(dev), camera_center=c.camera_center.to(dev))


def _close(got, want, what):
    r, d, s = common.float_err(got.detach().cpu().numpy(), want.detach().cpu().numpy())
    assert r <= 1.0, f"{what}: max|d|={d:.3e} max|ref|={s:.3e} tol-ratio={r:.2f}"


def _run(fn, model, loss_keys, dL, **kw):
    model.zero_grad()
    out = fn(**kw)
    loss = sum((out[k] * dL[k]).sum() for k in loss_keys)
    loss.backward()
    grads = {n: (None if t.grad is None else t.grad.clone()) for n, t in model.leaves().items()}
    grads["viewspace_points"] = out["viewspace_points"].grad.clone()
    return out, grads


CASES = [("render", dict(), ("render",)),
         ("render", dict(pipe_kw=dict(convert_SHs_python=True, compute_cov3D_python=True)), ("render",)),
         ("render", dict(filtered=True, scaling_modifier=0.8), ("render",)),
         ("render_mask", dict(), ("mask",)),
         ("render_with_depth", dict(), ("render", "mask")),
         ("render_with_depth", dict(filtered=True), ("render", "mask")),
         ("render_contrastive_feature", dict(call_kw=dict(norm_point_features=True)), ("render",))]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-{i}" for i, c in enumerate(CASES)])
def test_dropin_render_functions_match_the_reference_binding(case):
    name, opt, loss_keys = case
    ref, ours = _load_bindings()
    dev = torch.device("cuda", 0)
    P, H, W, K = 30000, 200, 304, 32
    sc = synthetic.scene(P, H, W, K)
    model, cam = _Model(sc, K, dev), _camera(sc, dev)
    pipe = SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    for k, v in opt.get("pipe_kw", {}).items():
        setattr(pipe, k, v)
    nch = K if name == "render_contrastive_feature" else 3
    bg = torch.linspace(0.1, 0.9, nch, device=dev)
    kw = dict(viewpoint_camera=cam, pc=model, pipe=pipe, bg_color=bg, **opt.get("call_kw", {}))
    if "scaling_modifier" in opt:
        kw["scaling_modifier"] = opt["scaling_modifier"]
    if opt.get("filtered"):
        kw["filtered_mask"] = (torch.arange(P, device=dev) % 7) == 0
    gen = torch.Generator().manual_seed(11)
    dL = {"render": (torch.randn(nch, H, W, generator=gen) / (H * W)).to(dev), "mask": (torch.randn(1 if name == "render_with_depth" else 3, H, W, generator=gen) / (H * W)).to(dev)}
    o_ref, g_ref = _run(getattr(ref, name), model, loss_keys, dL, **kw)
    o_our, g_our = _run(getattr(ours, name), model, loss_keys, dL, **kw)
    assert set(o_ref) == set(o_our)
    assert torch.equal(o_ref["radii"], o_our["radii"]) and torch.equal(o_ref["visibility_filter"], o_our["visibility_filter"])
    for k in o_ref:
        if k not in ("radii", "visibility_filter", "viewspace_points"):
            assert o_ref[k].shape == o_our[k].shape, k
            _close(o_our[k], o_ref[k], f"{name}[{k}]")
    for n in g_ref:
        assert (g_ref[n] is None) == (g_our[n] is None), n
        if g_ref[n] is not None:
            _close(g_our[n], g_ref[n], f"{name}: d/d{n}")
    assert g_ref["viewspace_points"].abs().max() > 0


def test_forward_mask_matches_the_reference_forward_mask():
    """DEPTH ``GaussianRasterizer.forward_mask`` (reference diff_gaussian_rasterization_depth/__init__.py:359-391): image and the
    gradient of the per-Gaussian mask against the reference's own mask-only kernels."""
    if not common.have_ref("depth"):
        pytest.skip("oracle/_ref not built")
    from seganygaussians_b200 import rasterizer as R
    refmod = common.ref_module("depth")
    dev = torch.device("cuda", 0)
    P, H, W = 20000, 160, 240
    sc = synthetic.scene(P, H, W, 3)
    g, c = sc.gauss, sc.cam
    gen = torch.Generator().manual_seed(3)
    dL = (torch.randn(1, H, W, generator=gen) / (H * W)).to(dev)
    res = {}
    for tag, Settings, Rast in (("ref", refmod.GaussianRasterizationSettings, refmod.GaussianRasterizer),
                                ("ours", R.GaussianRasterizationSettings, R.GaussianRasterizerDepth)):
        rs = Settings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
                      viewmatrix=c.world_view_transform.to(dev), projmatrix=c.full_proj_transform.to(dev), sh_degree=0,
                      campos=c.camera_center.to(dev), prefiltered=False, debug=False)
        mask = (torch.rand(P, 1, generator=torch.Generator().manual_seed(7)) * 0.5 + 0.5).to(dev).requires_grad_(True)
        out = Rast(raster_settings=rs).forward_mask(means3D=g.means3D.to(dev), means2D=torch.zeros(P, 3, device=dev), opacities=g.opacities.to(dev),
                                                    mask=mask, scales=g.scales.to(dev), rotations=g.rotations.to(dev), cov3D_precomp=None)
        img, radii = out[0], out[-1]
        (img * dL).sum().backward()
        res[tag] = (img.detach(), radii.detach(), mask.grad.detach())
    assert torch.equal(res["ref"][1], res["ours"][1])
    _close(res["ours"][0], res["ref"][0], "forward_mask image")
    _close(res["ours"][2], res["ref"][2], "forward_mask dL/dmask")
    assert res["ref"][2].abs().max() > 0
