#!/usr/bin/env python
"""BASELINE.json configs[3]: synthetic 3M Gaussians, 1080x1920, K=32, a batch of 8 cameras sharded over 1/2/4/8 GPUs with ONE
NCCL all-reduce of the per-Gaussian feature gradient per step (strong scaling: the batch is fixed, camera i -> rank i mod N).

    python tools/scale_c4.py --steps 10 --warmup 3                                   # 1 GPU: 8 cameras in sequence
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/scale_c4.py --steps 10 --warmup 3

Not the bench line (bench.py measures configs[1], weak scaling, as the contract asks); this is the measurement tool for the
configuration the north star quotes its ">= 6x at 8 GPUs" on.  Prints one JSON line on rank 0: Gaussian*pixel/s over the whole
batch, ms per batch (CUDA events, barrier + synchronize on both sides, max over ranks).  Written at the end of round 1 after
the GPU budget was spent: it has run on the CPU path of `render_camera_batch` (tests/test_dp_gloo.py) but not yet on a GPU."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gaussians", type=int, default=3_000_000)
    ap.add_argument("--cameras", type=int, default=8)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from seganygaussians_b200 import synthetic, rasterizer as R, _lib
    from seganygaussians_b200.data_parallel import render_camera_batch, FeatureGradReducer
    _lib.load()
    P, H, W, K = a.gaussians, 1080, 1920, 32
    g = synthetic.make_gaussians(P, K, W)
    cams = [synthetic.make_camera(H, W, i, n_cams=a.cameras) for i in range(a.cameras)]
    means3D, opac = g.means3D.to(dev), g.opacities.to(dev)
    scales, rots = g.scales.to(dev), g.rotations.to(dev)
    features = g.colors.to(dev).requires_grad_(True)
    means2D = torch.zeros(P, 3, device=dev)
    dL = synthetic.make_upstream(K, H, W).to(dev)
    bg = torch.zeros(K, device=dev)
    rasterizers = []
    for c in cams:
        rs = R.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg, scale_modifier=1.0,
                                             viewmatrix=c.world_view_transform.to(dev), projmatrix=c.full_proj_transform.to(dev),
                                             sh_degree=0, campos=c.camera_center.to(dev), prefiltered=False, debug=False)
        rasterizers.append(R.GaussianRasterizerContrastiveF(raster_settings=rs))

    def render_fn(cam_index, feats):
        color, _ = rasterizers[cam_index](means3D=means3D, means2D=means2D, opacities=opac, shs=None, colors_precomp=feats,
                                          scales=scales, rotations=rots, cov3D_precomp=None)
        return color

    reducer = FeatureGradReducer() if world > 1 else None

    def step():
        render_camera_batch(list(range(a.cameras)), render_fn, features, lambda img, ci: (img * dL).sum(), reducer=reducer,
                            rank=rank, world_size=world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        per_batch = float(ms.item()) / a.steps
        print(json.dumps({"metric": "fwd+bwd Gaussians*pixels/s @K=32", "value": float(P) * H * W * a.cameras / (per_batch * 1e-3),
                          "unit": "Gaussian*pixel/s", "n_gpus": world, "steps": a.steps, "ms_per_batch": per_batch, "scaling": "strong",
                          "config": {"workload": f"synthetic {P} Gaussians, {H}x{W}, K={K}, batch of {a.cameras} cameras sharded camera i -> rank i mod N, "
                                                 "one all-reduce of dL_dcolors per batch", "loss": "sum(image * dL) per camera (includes the loss kernels)"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
