"""Data parallelism for the training step: one process per GPU, rays sharded across ranks, parameters
replicated, ONE flat gradient buffer per module all-reduced (averaged) over NCCL/NVLink inside
`loss.backward()` -- so an unchanged trainer (train_hash2.py:226-228: backward, then optimiser steps) sees
reduced gradients.  The reference has no counterpart (it wraps only the MLP in nn.DataParallel); this is
SURVEY section 8(e).

Mechanics: `HashEncoder` / `MLP_3D` publish their flat gradient tensor (L,T,F) / (14227,) from inside their
autograd backward -- the encoder in level chunks, each as soon as its scatter-add kernel has been enqueued;
the hook below launches an asynchronous all-reduce on every published piece right there (NCCL's stream waits
for the producing kernel, the compute stream runs on: the MLP's reduce overlaps the hash-table backward, a
table chunk's reduce overlaps the next chunk's scatter-add) and registers an end-of-backward engine callback
that makes the compute stream wait for the collectives.  MSE is a mean over the LOCAL batch, so the
average over ranks equals the single-process gradient of the concatenated batch.
"""
from __future__ import annotations

import os
from typing import Iterable, List

import torch
import torch.distributed as dist
from torch.autograd import Variable


def init_from_env(backend: str | None = None) -> tuple[int, int]:
    """Initialises torch.distributed from RANK/WORLD_SIZE/MASTER_* (torchrun) and pins the CUDA device.
    Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


class GradAllReduce:
    """Attach to native modules: `GradAllReduce(encoder, mlp)`.  Detach with `.remove()`."""

    def __init__(self, *modules, group=None, average: bool = True):
        self.group = group
        self.average = average
        self.modules = list(modules)
        self._pending: List = []
        self._callback_queued = False
        self.bytes_reduced = 0
        for m in self.modules:
            m._grad_hooks.append(self._on_grad)

    def remove(self):
        for m in self.modules:
            if self._on_grad in m._grad_hooks:
                m._grad_hooks.remove(self._on_grad)

    def _on_grad(self, flat: torch.Tensor):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        op = dist.ReduceOp.SUM
        if self.average:
            if dist.get_backend(self.group) == "nccl":
                op = dist.ReduceOp.AVG                     # averaging inside the collective: no extra pass over the buffer
            else:
                flat.div_(dist.get_world_size(self.group))
        self._pending.append(dist.all_reduce(flat, op=op, group=self.group, async_op=True))
        self.bytes_reduced += flat.numel() * flat.element_size()
        if not self._callback_queued:
            self._callback_queued = True
            Variable._execution_engine.queue_callback(self._finish)

    def _finish(self):
        for w in self._pending:
            w.wait()                       # compute stream waits on the NCCL stream; no host sync for CUDA tensors
        self._pending.clear()
        self._callback_queued = False


def shard_rays(n_rays: int, rank: int, world: int) -> slice:
    """Rank r takes rays [r*n/world, (r+1)*n/world)."""
    per = n_rays // world
    return slice(rank * per, (rank + 1) * per if rank < world - 1 else n_rays)


def slab_range(res: int, rank: int, world: int) -> tuple[int, int]:
    """Planes [i0, i1) of axis 0 of the density grid owned by `rank` (nerf2mesh grid sharding)."""
    base, rem = divmod(res, world)
    i0 = rank * base + min(rank, rem)
    return i0, i0 + base + (1 if rank < rem else 0)


def allgather_counts(counts: Iterable[int]) -> list[list[int]]:
    """Per-rank (vertices, triangles) -> list over ranks; exclusive prefix gives each rank's output offsets."""
    t = torch.tensor(list(counts), dtype=torch.int64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [t.tolist()]
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]
