"""CPU, world_size = 2, gloo: the host logic of the image-batch data parallelism (camera sharding, local
accumulation, ONE all-reduce of the per-Gaussian feature gradient).  The render call is injected, so here the CPU
oracle stands in for the CUDA rasterizer (tests may do that; the product path never does)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import common
from seganygaussians_b200 import synthetic
from seganygaussians_b200.data_parallel import shard_cameras, render_camera_batch, FeatureGradReducer


def test_shard_cameras_partition():
    for n in (0, 1, 7, 8, 9):
        for w in (1, 2, 4, 8):
            shards = [shard_cameras(n, r, w) for r in range(w)]
            assert sorted(sum(shards, [])) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        shard_cameras(4, 2, 2)


class _OracleRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, scene_args):
        from oracle import oracle
        P, H, W, K, cam = scene_args
        sc = synthetic.scene(P, H, W, K, cam=cam)
        g, c = sc.gauss, sc.cam
        fw = oracle.forward(means3D=g.means3D.numpy(), opacities=g.opacities.numpy(), bg=np.zeros(K, np.float32),
                            viewmatrix=c.world_view_transform.numpy(), projmatrix=c.full_proj_transform.numpy(),
                            campos=c.camera_center.numpy(), image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy,
                            colors_precomp=features.detach().numpy(), scales=g.scales.numpy(), rotations=g.rotations.numpy())
        ctx.fw = fw
        return torch.from_numpy(fw.color.copy())

    @staticmethod
    def backward(ctx, grad):
        from oracle import oracle
        bw = oracle.backward(ctx.fw, grad.contiguous().numpy())
        return torch.from_numpy(bw.colors.copy()), None


def _worker(rank, world, port, n_cams, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, H, W, K = 400, 32, 48, 8
    feats = synthetic.make_gaussians(P, K, W).colors.clone().requires_grad_(True)
    cams = list(range(n_cams))
    dLs = [synthetic.make_upstream(K, H, W, seed=10 + i) for i in cams]
    calls = []
    if n_cams == 4:      # one of the two cases goes through a caller-supplied exchange (what bench.py does with the multicast all-reduce)
        reducer = FeatureGradReducer(side_stream=False, reduce_fn=lambda g: (calls.append(tuple(g.shape)), dist.all_reduce(g))[1])
    else:
        reducer = FeatureGradReducer(side_stream=False)
    loss, mine = render_camera_batch(cams, lambda cam, f: _OracleRender.apply(f, (P, H, W, K, cam)), feats,
                                     lambda img, ci: (img * dLs[ci]).sum(), reducer=reducer)
    assert calls == ([(P, K)] if n_cams == 4 else [])
    out[rank] = (feats.grad.clone(), mine, loss)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_cams", [3, 4])
def test_two_rank_feature_grad_allreduce_matches_single_process(n_cams):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_cams, out), nprocs=world, join=True)
    g0, mine0, _ = out[0]
    g1, mine1, _ = out[1]
    assert sorted(mine0 + mine1) == list(range(n_cams)) and mine0 == list(range(0, n_cams, 2))
    assert torch.equal(g0, g1)                       # every rank holds the same reduced gradient
    # single-process reference: all cameras on one rank, no collective
    P, H, W, K = 400, 32, 48, 8
    feats = synthetic.make_gaussians(P, K, W).colors.clone().requires_grad_(True)
    dLs = [synthetic.make_upstream(K, H, W, seed=10 + i) for i in range(n_cams)]
    render_camera_batch(list(range(n_cams)), lambda cam, f: _OracleRender.apply(f, (P, H, W, K, cam)), feats,
                        lambda img, ci: (img * dLs[ci]).sum(), rank=0, world_size=1)
    assert torch.allclose(g0, feats.grad, rtol=1e-5, atol=1e-9)
