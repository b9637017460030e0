"""Data parallelism for the training step: one process per GPU, rays sharded across ranks, parameters
replicated, ONE flat gradient buffer per module all-reduced (averaged) over NCCL/NVLink inside
`loss.backward()` -- so an unchanged trainer (train_hash2.py:226-228: backward, then optimiser steps) sees
reduced gradients.  The reference has no counterpart (it wraps only the MLP in nn.DataParallel); this is
SURVEY section 8(e).

Two exchanges are provided: `GradAllReduce` (NCCL calls, below) and `PeerGradAllReduce` (ONE kernel of this
package over NVLink peer memory / NVLS multicast, csrc/comm.cu; the default of `attach_grad_allreduce`).

Mechanics of the NCCL variant: `HashEncoder` / `MLP_3D` publish their flat gradient tensor (L,T,F) / (14227,) from inside their
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


class PeerGradAllReduce:
    """`PeerGradAllReduce(encoder, mlp)`: the same contract as GradAllReduce, but the exchange is ONE kernel over NVLink
    peer memory (csrc/comm.cu) instead of NCCL calls.  The two modules' flat gradients become slices of one persistent
    peer-mapped region [table (L,T,F) | MLP parameters]; the scatter-add of the hash-grid backward and the MLP
    gradient reduction write straight into it, and an end-of-backward engine callback enqueues hbr_allreduce_peer on
    the compute stream (no extra stream, no event round trip, capturable in the step's CUDA graph).  Every rank ends
    with bit-identical averaged gradients.  `.grad` of the parameters aliases the region from step to step (use
    zero_grad(set_to_none=True), torch's default)."""

    def __init__(self, encoder, mlp, group=None, average: bool = True, transport: str = "ipc", ctas: int = 0,
                 overlap: bool = False, chunks: int = 2):
        """overlap=False: one all-reduce of the whole region behind the backward pass.  overlap=True: the table backward
        runs in `chunks` level chunks and every published piece (the MLP gradient first, then each level chunk) is
        all-reduced at once by a small grid (`ctas`, default 32) on a high-priority side stream while the next chunk's
        scatter-add still runs on the compute stream."""
        from .peer import PeerRegion
        self.group, self.average, self.overlap = group, average, overlap
        self.ctas = ctas if ctas > 0 else (32 if overlap else 0)
        self.modules = [encoder, mlp]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n_tab = encoder.L * encoder.T * encoder.F
        n_mlp = mlp._flat_params().numel()
        off = -(-n_tab // 4) * 4
        self.region = PeerRegion(off + n_mlp, group=group, transport=transport)
        encoder._flat_table()
        encoder._grad_buffer = self.region.tensor[:n_tab].view(encoder.L, encoder.T, encoder.F)
        mlp._grad_buffer = self.region.tensor[off:off + n_mlp]
        if overlap:
            encoder._grad_chunks = chunks
        self._side = torch.cuda.Stream(device=self.region.device, priority=-1) if overlap else None
        self._base = self.region.tensor.data_ptr()
        self._callback_queued = False
        self._side_used = False
        self.bytes_reduced = 0
        for m in self.modules:
            m._grad_hooks.append(self._on_grad)

    def remove(self):
        for m in self.modules:
            if self._on_grad in m._grad_hooks:
                m._grad_hooks.remove(self._on_grad)
            m._grad_buffer = None

    def _scale(self):
        return 1.0 / self.world if self.average else 1.0

    def _on_grad(self, flat: torch.Tensor):
        if not self._callback_queued:
            self._callback_queued = True
            Variable._execution_engine.queue_callback(self._finish)
        if self.overlap:
            off = (flat.data_ptr() - self._base) // 4
            n = -(-flat.numel() // 4) * 4
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)                       # the piece's producer kernel has been enqueued on `cur`
            with torch.cuda.stream(self._side):
                self.region.all_reduce(n=n, scale=self._scale(), ctas=self.ctas, offset=off)
            self._side_used = True
            self.bytes_reduced += n * 4

    def _finish(self):
        self._callback_queued = False
        if self.overlap:
            if self._side_used:
                torch.cuda.current_stream().wait_stream(self._side)
                self._side_used = False
            return
        self.region.all_reduce(scale=self._scale(), ctas=self.ctas)
        self.bytes_reduced += self.region.n * 4


def attach_grad_allreduce(encoder, mlp, group=None, kind: str = "auto", **peer_kw):
    """The gradient exchange for `world` > 1: kind "peer" = PeerGradAllReduce, "nccl" = GradAllReduce, "auto" = peer
    when every rank can set it up (CUDA ranks of one node), else NCCL.  Peer transport "auto": 2 ranks exchange through
    plain peer loads/stores over CUDA-IPC mappings (measured 0.130 ms for 67 MB against 0.153 ms NCCL and 0.21 ms
    multicast); more ranks reduce inside the NVSwitch through an NVLS multicast mapping when torch's symmetric memory
    offers one (8 ranks: 0.184 ms against 0.240 ms peer loads/stores and 0.245 ms NCCL)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    if kind == "nccl" or not torch.cuda.is_available() or dist.get_backend(group) != "nccl":
        return GradAllReduce(encoder, mlp, group=group)
    world = dist.get_world_size(group)
    kw = dict(peer_kw)
    kw.setdefault("overlap", True)        # measured best at 4 096 rays/GPU: 2 level chunks, 32 CTAs on the side stream
    kw.setdefault("chunks", 2)            # (W=8: 0.656 ms/step against 0.712 behind the backward, 0.671 with 4 chunks)
    if kw.get("transport", "auto") == "auto":
        kw["transport"] = "ipc" if world == 2 else "symm"
    red, err = None, None
    try:
        red = PeerGradAllReduce(encoder, mlp, group=group, **kw)
    except Exception as e:                                   # e.g. no peer access between the devices
        err = e
    ok = torch.tensor([1 if red is not None else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 1:
        return red
    if red is not None:
        red.remove()
    if kind == "peer":
        raise RuntimeError(f"peer-memory all-reduce unavailable on at least one rank: {err!r}")
    return GradAllReduce(encoder, mlp, group=group)


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
