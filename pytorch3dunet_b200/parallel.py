"""Data-parallel replicas: one process per GPU, the gradients are the only thing exchanged.

Replaces the reference's single-process `nn.DataParallel` (trainer.py:203-204: per-step parameter broadcast,
scatter/gather through GPU 0, gradient reduce-to-GPU-0) with identical replicas.  Two reducers:

* `optim.BucketedAllReduce` over `optim.FlatParameters` (what bench.py uses): the engine writes gradients straight into one flat
  buffer and each bucket's NCCL allreduce starts while backward is still running;
* `GradAllReducer` below: the plain fallback for arbitrary parameter lists (flatten, one all_reduce, average, scatter back) --
  gloo in the CPU tests.

Semantics note (differs from the reference and is deliberate): every replica computes the loss of ITS OWN per-GPU batch and the
gradients are averaged.  nn.DataParallel gathers the outputs and evaluates the loss once on the global batch; for the mean-reduced
BCE term the two agree, for the Dice term (a ratio of batch sums, losses.py:11-37) they do not -- averaging per-replica Dice
gradients is what torch DistributedDataParallel would do as well.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int):
    """patch i -> rank i % world (the reference multiplies the loader batch by the GPU count instead,
    datasets/utils.py:399-403); used for inference patch sharding (no collective needed)."""
    return list(range(rank, n_items, world))


class GradAllReducer:
    def __init__(self, params, world: int | None = None):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if world is None else world
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def __call__(self):
        if self.world == 1 or not self.params:
            return
        dev = next((p.grad.device for p in self.params if p.grad is not None), self.params[0].device)
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.empty(self.numel, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:  # unused / frozen on this rank this step: contributes zeros (all ranks reduce the same layout)
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(self.flat)
        self.flat.div_(self.world)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = self.flat[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(self.flat[off:off + n].view_as(p.grad))
            off += n
