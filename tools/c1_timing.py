#!/usr/bin/env python
"""Developer diagnostic: fwd / bwd CUDA-event times at BASELINE configs[0] size (10k Gaussians, 256x256, K=3), ours and -- when
oracle/_ref is present -- the unmodified reference.  Not a bench."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seganygaussians_b200 import synthetic, rasterizer as R
from tests import common

def run(mod_settings, mod_rast, sc, dev, iters=200):
    g, c = sc.gauss, sc.cam
    P, K = sc.P, 3
    t = [x.to(dev).requires_grad_(True) for x in (g.means3D, torch.zeros(P, 3), g.opacities, g.scales, g.rotations, g.colors[:, :3].contiguous())]
    rs = mod_settings(image_height=sc.H, image_width=sc.W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
                      viewmatrix=c.world_view_transform.to(dev), projmatrix=c.full_proj_transform.to(dev), sh_degree=0,
                      campos=c.camera_center.to(dev), prefiltered=False, debug=False)
    rast = mod_rast(raster_settings=rs)
    dL = sc.dL_dout[:3].to(dev)
    def step():
        for x in t: x.grad = None
        out = rast(means3D=t[0], means2D=t[1], opacities=t[2], shs=None, colors_precomp=t[5], scales=t[3], rotations=t[4], cov3D_precomp=None)
        return out[0]
    for _ in range(10): step().backward(dL)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(iters):
        e[0].record(); c_ = step(); e[1].record(); c_.backward(dL); e[2].record(); torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    return tf / iters, tb / iters

dev = torch.device("cuda", 0)
sc = synthetic.scene(10000, 256, 256, 3)
for binning in ("depth_first", "radix"):
    R.set_binning(binning)
    print("ours [%s]: fwd %.3f ms  bwd %.3f ms" % ((binning,) + run(R.GaussianRasterizationSettings, R.GaussianRasterizer, sc, dev)), flush=True)
R.set_binning()
if common.have_ref("base"):
    ref = common.ref_module("base")
    print("reference: fwd %.3f ms  bwd %.3f ms" % run(ref.GaussianRasterizationSettings, ref.GaussianRasterizer, sc, dev), flush=True)
