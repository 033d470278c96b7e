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

    def __init__(self, group=None, side_stream: bool = True, reduce_fn=None):
        self.group = group
        self.reduce_fn = reduce_fn          # e.g. MulticastAllReduce.all_reduce_ (the library's own exchange); None: NCCL
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stream = None
        self.side_stream = side_stream
        self._pending = None

    def reduce_async(self, grad: torch.Tensor):
        if self.world == 1:
            return None
        if grad.is_cuda and self.side_stream:
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=grad.device, priority=-1)   # ahead of queued blocks of the main stream
            self.stream.wait_stream(torch.cuda.current_stream(grad.device))
            with torch.cuda.stream(self.stream):
                if self.reduce_fn is not None:
                    self.reduce_fn(grad)
                else:
                    dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
            grad.record_stream(self.stream)
            self._pending = ("stream", grad.device)
        elif self.reduce_fn is not None:
            self.reduce_fn(grad)                   # the caller's own exchange, synchronous on this path
            self._pending = None
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


class MulticastAllReduce:
    """The library's own all-reduce (sum, fp32) over the NVSwitch multicast mapping: ``sagars_multimem_allreduce_f32``
    (csrc/multimem_allreduce.cu) on a symmetric buffer from ``torch.distributed._symmetric_memory``.  Opt-in alternative to the
    NCCL call (``bench.py --allreduce multimem``); needs NVSwitch multicast support (``multicast_ptr != 0``), otherwise the
    constructor raises and the caller stays with NCCL.

    ``all_reduce_(t)``: copy ``t`` into the symmetric buffer, barrier (every rank's copy is complete), one pass in which rank r
    reduces the r-th slice inside the switch and broadcasts it, barrier (every slice has landed everywhere), copy back -- all on
    the current stream.  The two copies cost 2 x numel x 8 B of local HBM traffic (0.08 ms for the 128 MB tensor of the headline
    configuration); a trainer that lets the rasterizer accumulate straight into ``self.buffer`` avoids them."""

    def __init__(self, numel: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        if numel % 4:
            raise ValueError("numel must be a multiple of 4 (16-byte multimem accesses)")
        self._lib = _lib.load()
        self._check = _lib.check
        self.group = group if group is not None else dist.group.WORLD
        self.device = torch.device(device)
        self.numel = int(numel)
        self.buffer = symm.empty(self.numel, dtype=torch.float32, device=self.device)
        self.handle = symm.rendezvous(self.buffer, self.group)
        self.multicast_ptr = int(self.handle.multicast_ptr)
        if self.multicast_ptr == 0:
            raise RuntimeError("this device / fabric exposes no multicast mapping for symmetric memory: stay with NCCL")
        self.rank, self.world = int(self.handle.rank), int(self.handle.world_size)

    def all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != self.numel or t.device != self.buffer.device:
            raise ValueError("all_reduce_: expected a contiguous float32 tensor of the size this object was created for")
        stream = torch.cuda.current_stream(self.buffer.device)
        self.buffer.copy_(t.view(-1))
        self.handle.barrier(channel=0)
        dev = self.buffer.device.index if self.buffer.device.index is not None else torch.cuda.current_device()
        self._check(self._lib.sagars_multimem_allreduce_f32(dev, self.multicast_ptr, self.numel, self.rank, self.world,
                                                            int(stream.cuda_stream)))
        self.handle.barrier(channel=1)
        t.view(-1).copy_(self.buffer)
        return t


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
