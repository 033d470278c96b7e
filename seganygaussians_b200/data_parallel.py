"""Image-batch data parallelism for the rasterizer path: one process per GPU, cameras sharded across
ranks, Gaussians replicated, ONE collective per step -- the sum of the per-Gaussian feature gradients
``dL_dcolors [P, K]`` (the only tensor SAGA's contrastive training optimises,
``scene/gaussian_model_ff.py:154-162`` of the reference).  The reference itself is single-GPU
(``utils/general_utils.py:133``); this module is what the north star adds.

The render call is injected (``render_fn``), so the host logic (sharding, accumulation, collective) is
exercised on CPU with the ``gloo`` backend in tests, and with NCCL over NVLink on the GPU box.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_cameras(num_cameras: int, rank: int, world_size: int) -> List[int]:
    """Camera i goes to rank i mod world_size."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    return list(range(rank, num_cameras, world_size))


class FeatureGradReducer:
    """Sums a per-Gaussian gradient tensor across ranks, optionally on a side stream so the collective
    overlaps whatever the caller does next (the per-Gaussian geometry backward does not touch
    ``dL_dcolors`` when colours are precomputed)."""

    def __init__(self, group=None, side_stream: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stream = None
        self.side_stream = side_stream
        self._pending = None

    def reduce_async(self, grad: torch.Tensor):
        if self.world == 1:
            return None
        if grad.is_cuda and self.side_stream:
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=grad.device)
            self.stream.wait_stream(torch.cuda.current_stream(grad.device))
            with torch.cuda.stream(self.stream):
                dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
            grad.record_stream(self.stream)
            self._pending = ("stream", grad.device)
        else:
            self._pending = ("work", dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        return self._pending

    def ready_event(self):
        """A CUDA event recorded behind the pending all-reduce on its stream (None when nothing is pending or the reduction
        does not run on a side stream).  Hand it to ``rasterizer.set_blend_wait_event``: the next forward's geometry stages
        then overlap the exchange and only its blend stage -- the first reader of the features -- waits for it."""
        if self._pending is None or self._pending[0] != "stream":
            return None
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return ev

    def wait(self):
        if self._pending is None:
            return
        kind, h = self._pending
        if kind == "stream":
            torch.cuda.current_stream(h).wait_stream(self.stream)
        else:
            h.wait()
        self._pending = None


def render_camera_batch(cameras: Sequence, render_fn: Callable, features: torch.Tensor,
                        loss_fn: Callable[[torch.Tensor, int], torch.Tensor],
                        reducer: Optional[FeatureGradReducer] = None, rank: Optional[int] = None,
                        world_size: Optional[int] = None):
    """One data-parallel step over a batch of cameras.

    Every rank renders the cameras ``shard_cameras(len(cameras), rank, world)``, back-propagates
    ``loss_fn(image, camera_index)`` and accumulates into ``features.grad`` locally; then ONE all-reduce
    sums ``features.grad`` over ranks.  Returns (local loss sum, list of local camera indices).
    """
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    mine = shard_cameras(len(cameras), rank, world_size)
    if features.grad is not None:
        features.grad = None
    total = 0.0
    for ci in mine:
        image = render_fn(cameras[ci], features)
        loss = loss_fn(image, ci)
        loss.backward()
        total = total + float(loss.detach())
    if features.grad is None:   # a rank with no camera still has to take part in the collective
        features.grad = torch.zeros_like(features)
    if reducer is None:
        reducer = FeatureGradReducer()
    reducer.reduce_async(features.grad)
    reducer.wait()
    return total, mine
