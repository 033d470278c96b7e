#!/usr/bin/env python
"""Developer diagnostic: kernel timeline of one resident fwd+bwd step at the bench workload (torch.profiler / CUPTI),
to see where the GPU idles between our launches.  Not a bench: numbers under a profiler are never reported."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seganygaussians_b200 import synthetic, rasterizer as R, _lib  # noqa: E402


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

    e2e = "--e2e" in sys.argv
    cam_h = torch.cat([c.world_view_transform.reshape(-1), c.full_proj_transform.reshape(-1), torch.zeros(K),
                       c.camera_center.reshape(-1)]).float().pin_memory()      # bench.py's packed camera buffer

    def step():
        for t in leaves:
            t.grad = None
        t0 = time.perf_counter()
        if e2e:   # bench.py's end-to-end step: camera from pinned host memory, loss scalar read back
            cam = cam_h.to(dev, non_blocking=True)
            r = R.GaussianRasterizerContrastiveF(raster_settings=rs._replace(
                viewmatrix=cam[0:16].view(4, 4), projmatrix=cam[16:32].view(4, 4), bg=cam[32:32 + K], campos=cam[32 + K:35 + K]))
        else:
            r = rast
        color, radii = r(means3D=means3D, means2D=means2D, opacities=opac, shs=None, colors_precomp=colors,
                         scales=scales, rotations=rots, cov3D_precomp=None)
        t1 = time.perf_counter()
        if e2e:
            loss = torch.dot(color.reshape(-1), dL.reshape(-1))
            loss.backward()
            float(loss.item())
        else:
            color.backward(dL)
        t2 = time.perf_counter()
        marks.append((t0, t1, t2))

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    marks.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"plain loop: {e0.elapsed_time(e1) / 20:.3f} ms/step")
    f = [(b - a_) * 1e3 for a_, b, _ in marks]
    bw = [(c_ - b) * 1e3 for _, b, c_ in marks]
    per = [(marks[i + 1][0] - marks[i][0]) * 1e3 for i in range(len(marks) - 1)]
    print(f"host: forward call {sum(f) / len(f):.3f} ms, backward call {sum(bw) / len(bw):.3f} ms, step period {sum(per) / len(per):.3f} ms")

    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print("no CUDA events captured")
        return
    # the middle step: between the 2nd and 3rd preprocess kernel
    starts = [i for i, e in enumerate(evs) if "preprocess_kernel" in e.name]
    lo, hi = (starts[1], starts[2]) if len(starts) >= 3 else (0, len(evs))
    t_base = evs[lo].time_range.start
    prev_end = t_base
    busy = 0.0
    print(f"{'start_us':>10} {'dur_us':>9} {'gap_us':>8}  kernel")
    for e in evs[lo:hi]:
        s, d = e.time_range.start - t_base, e.time_range.end - e.time_range.start
        gap = e.time_range.start - prev_end
        busy += d
        print(f"{s:10.1f} {d:9.1f} {gap:8.1f}  {e.name[:90]}")
        prev_end = max(prev_end, e.time_range.end)
    span = evs[hi].time_range.start - t_base if hi < len(evs) else prev_end - t_base
    print(f"step span {span:.1f} us, busy {busy:.1f} us, idle {span - busy:.1f} us")


if __name__ == "__main__":
    main()
