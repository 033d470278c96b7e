"""Checks of the multi-GPU pieces (the all-reduce test needs a box with >= 2 GPUs: ``gpurun --gpus 2 -- 'python -m pytest
tests/test_multi_gpu.py -m gpu -q'``; on one GPU it is skipped, the event-gating test runs everywhere).

  * the library's own all-reduce over the NVSwitch multicast mapping against NCCL's result on the same tensor;
  * a forward whose blend stage is gated on an event (ABI v3 ``blend_wait_event``) gives the same image as an ungated one, and
    really waits: the event is recorded behind a long-running kernel on another stream."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _allreduce_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from seganygaussians_b200.data_parallel import MulticastAllReduce
        n = 1_000_000 * 32
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        t = torch.randn(n, generator=g).to(dev)
        ref = t.clone()
        dist.all_reduce(ref)
        own = MulticastAllReduce(n, dev)
        got = own.all_reduce_(t.clone())
        got2 = own.all_reduce_(t.clone())            # the object is reusable
        torch.cuda.synchronize(dev)
        err = float((got - ref).abs().max() / ref.abs().max())
        out[rank] = (err, bool(torch.equal(got, got2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multicast_allreduce_matches_nccl():
    import torch.multiprocessing as mp
    world = 2
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_allreduce_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        err, reproducible = out[rank]
        assert err < 1e-6, (rank, err)            # same two addends per element at world 2; order only matters from 3 ranks on
        assert reproducible


def test_blend_wait_event_gates_only_the_blend():
    from tests import common
    from seganygaussians_b200 import synthetic, rasterizer as R
    dev = torch.device("cuda", 0)
    P, H, W, K = 20000, 270, 480, 32
    sc = synthetic.scene(P, H, W, K)
    g, c = sc.gauss, sc.cam
    rs = R.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=torch.zeros(K, device=dev),
                                         scale_modifier=1.0, viewmatrix=c.world_view_transform.to(dev), projmatrix=c.full_proj_transform.to(dev),
                                         sh_degree=0, campos=c.camera_center.to(dev), prefiltered=False, debug=False)
    rast = R.GaussianRasterizerContrastiveF(raster_settings=rs)
    args = dict(means3D=g.means3D.to(dev), means2D=torch.zeros(P, 3, device=dev), opacities=g.opacities.to(dev), shs=None,
                scales=g.scales.to(dev), rotations=g.rotations.to(dev), cov3D_precomp=None)
    feats = g.colors.to(dev)
    with torch.no_grad():
        plain, _ = rast(colors_precomp=feats, **args)
        # the "optimiser" on a side stream: a long kernel, then the features change, then the event
        side = torch.cuda.Stream(device=dev)
        feats2 = feats.clone()
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            torch.cuda._sleep(200_000_000)                 # ~0.1 s of GPU time
            feats2.mul_(-1.0)                              # what the blend must see
            ev = torch.cuda.Event()
            ev.record(side)
        R.set_blend_wait_event(ev)
        gated, _ = rast(colors_precomp=feats2, **args)     # queued at once; only its blend waits for the event
        torch.cuda.synchronize(dev)
    assert torch.allclose(gated, -plain, rtol=1e-5, atol=1e-7)
