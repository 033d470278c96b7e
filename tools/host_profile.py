#!/usr/bin/env python
"""Developer diagnostic: where does HOST time go in one resident fwd+bwd step (bench workload)?  Wall-clock phases of
the step, then a cProfile of 30 steps.  Not a bench."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seganygaussians_b200 import synthetic, rasterizer as R  # noqa: E402


def main():
    P, H, W, K = 1_000_000, 1080, 1920, 32
    dev = torch.device("cuda", 0)
    sc = synthetic.scene(P, H, W, K)
    g, c = sc.gauss, sc.cam
    leaves = [t.to(dev).requires_grad_(True) for t in (g.means3D, torch.zeros(P, 3), g.opacities, g.scales, g.rotations, g.colors)]
    means3D, means2D, opac, scales, rots, colors = leaves
    dL = sc.dL_dout.to(dev)
    rs = R.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy,
                                         bg=torch.zeros(K, device=dev), scale_modifier=1.0,
                                         viewmatrix=c.world_view_transform.to(dev), projmatrix=c.full_proj_transform.to(dev),
                                         sh_degree=0, campos=c.camera_center.to(dev), prefiltered=False, debug=False)
    rast = R.GaussianRasterizerContrastiveF(raster_settings=rs)
    marks = []

    def step():
        t0 = time.perf_counter()
        for t in leaves:
            t.grad = None
        t1 = time.perf_counter()
        color, radii = rast(means3D=means3D, means2D=means2D, opacities=opac, shs=None, colors_precomp=colors,
                            scales=scales, rotations=rots, cov3D_precomp=None)
        t2 = time.perf_counter()
        color.backward(dL)
        t3 = time.perf_counter()
        marks.append((t0, t1, t2, t3))

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    marks.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        step()
    e1.record()
    torch.cuda.synchronize()
    n = len(marks)
    med = lambda xs: sorted(xs)[len(xs) // 2]
    print(f"device loop {e0.elapsed_time(e1) / n:.3f} ms/step; host medians (ms): grad=None {med([(b - a) * 1e3 for a, b, _, _ in marks]):.3f}  "
          f"forward {med([(c_ - b) * 1e3 for _, b, c_, _ in marks]):.3f}  backward {med([(d - c_) * 1e3 for _, _, c_, d in marks]):.3f}  "
          f"period {med([(marks[i + 1][0] - marks[i][0]) * 1e3 for i in range(n - 1)]):.3f}")
    print("cpu count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "loadavg", os.getloadavg())
    try:
        print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
    except OSError:
        pass

    # same loop with the GPU idle at every step start (sync) -> pure host cost of the calls
    hf, hb = [], []
    for _ in range(10):
        torch.cuda.synchronize()
        for t in leaves:
            t.grad = None
        a = time.perf_counter()
        color, radii = rast(means3D=means3D, means2D=means2D, opacities=opac, shs=None, colors_precomp=colors,
                            scales=scales, rotations=rots, cov3D_precomp=None)
        b = time.perf_counter()
        color.backward(dL)
        d = time.perf_counter()
        hf.append((b - a) * 1e3); hb.append((d - b) * 1e3)
    print(f"idle-GPU call cost (ms): forward(incl. count read-back) {med(hf):.3f}  backward(launch only) {med(hb):.3f}")

    pr = cProfile.Profile()
    pr.enable()
    for _ in range(30):
        step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
