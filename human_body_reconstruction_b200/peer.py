"""Peer-mapped gradient region + the one-kernel all-reduce over it (csrc/comm.cu).

`PeerRegion(n)` is n fp32 of device memory (plus the barrier flag words behind them) that every rank of the process
group has mapped: transport "ipc" = a plain cudaMalloc exported through a CUDA IPC handle (hbr_peer_export/import, the
64-byte handles travel through torch.distributed.all_gather_object); transport "symm" = torch's symmetric-memory
allocator, which additionally offers an NVLS multicast mapping when driver and fabric support it (then the reduction
runs inside the NVSwitch).  `all_reduce()` enqueues hbr_allreduce_peer on the current stream.  torch is plumbing here
(rendezvous, tensor views); the data path is the CUDA kernel.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib


class _Raw:
    """__cuda_array_interface__ carrier: lets torch view memory this package allocated itself."""

    def __init__(self, ptr: int, n: int, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
        self._owner = owner


def _round_up(a: int, b: int) -> int:
    return -(-a // b) * b


class PeerRegion:
    def __init__(self, n_floats: int, group=None, transport: str = "ipc", multicast: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("PeerRegion needs CUDA devices (there is no CPU path)")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > _lib.HBR_MAX_PEERS:
            raise ValueError(f"peer all-reduce is built for <= {_lib.HBR_MAX_PEERS} ranks of one NVSwitch domain")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.n = _round_up(int(n_floats), 4 * self.world)                 # whole float4 slices per rank
        self._flag_off = _round_up(self.n * 4, 256)
        self._status_off = self._flag_off + _lib.HBR_PEER_FLAG_BYTES
        self.bytes = self._status_off + 256
        self.transport = transport
        self.multicast_ptr = 0
        self._own = None
        self._imported = []
        self._symm = None
        if transport == "ipc":
            self._init_ipc()
        elif transport == "symm":
            self._init_symm(multicast)
        else:
            raise ValueError(transport)
        W = self.world
        self._bufs = (C.c_void_p * W)(*[C.c_void_p(p) for p in self.ptrs])
        self._flags = (C.c_void_p * W)(*[C.c_void_p(p + self._flag_off) for p in self.ptrs])
        self._status_ptr = C.c_void_p(self.ptrs[self.rank] + self._status_off)
        self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.calls = 0

    # -- transports ------------------------------------------------------------------------------------------
    def _init_ipc(self):
        own = C.c_void_p()
        check(lib().hbr_peer_alloc(C.byref(own), self.bytes))
        self._own = own
        handle = (C.c_ubyte * _lib.HBR_PEER_HANDLE_BYTES)()
        check(lib().hbr_peer_export(own, handle))
        handles = [bytes(handle)]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=self.group)
        self.ptrs = []
        for q, h in enumerate(handles):
            if q == self.rank:
                self.ptrs.append(own.value)
                continue
            p = C.c_void_p()
            check(lib().hbr_peer_import((C.c_ubyte * _lib.HBR_PEER_HANDLE_BYTES)(*h), C.byref(p)))
            self._imported.append(p)
            self.ptrs.append(p.value)
        self._whole = torch.as_tensor(_Raw(own.value, self.bytes // 4, self), device=self.device)
        self.tensor = self._whole[: self.n]
        if self.world > 1:
            dist.barrier(group=self.group)

    def _init_symm(self, multicast: bool):
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        t = symm_mem.empty(self.bytes // 4, dtype=torch.float32, device=self.device)
        t.zero_()
        torch.cuda.synchronize()
        hdl = symm_mem.rendezvous(t, grp)
        self._symm = (t, hdl)
        self.ptrs = [int(p) for p in hdl.buffer_ptrs]
        if multicast and getattr(hdl, "has_multicast_support", None) is not None:
            try:
                self.multicast_ptr = int(hdl.multicast_ptr or 0)
            except Exception:
                self.multicast_ptr = 0
        self._whole = t
        self.tensor = t[: self.n]
        dist.barrier(group=self.group)

    # -- the collective --------------------------------------------------------------------------------------
    def all_reduce(self, n: Optional[int] = None, scale: float = 1.0, ctas: int = 0, offset: int = 0):
        """In place over floats [offset, offset + n) (default: the whole region): tensor <- scale * sum over ranks.
        Enqueued on the current stream; every rank must call it with the same offset / n / ctas, and calls that may run
        concurrently (different streams) are not allowed -- they share the barrier flags."""
        n = self.n - offset if n is None else _round_up(int(n), 4)
        if offset % 4 or offset < 0 or offset + n > self.n:
            raise ValueError("offset must be a multiple of 4 floats and the range must lie inside the region")
        bufs = self._bufs
        if offset:
            bufs = (C.c_void_p * self.world)(*[C.c_void_p(p + 4 * offset) for p in self.ptrs])
        mc = C.c_void_p(self.multicast_ptr + 4 * offset) if self.multicast_ptr else None
        check(lib().hbr_allreduce_peer(bufs, self._flags, mc, self.rank, self.world, n, float(scale), int(ctas),
                                       self._status_ptr, _lib.stream()))
        self.calls += 1

    def all_reduce_stream(self, pieces, done: Optional[torch.Tensor], scale: float = 1.0, ctas: int = 0):
        """The exchange as ONE launch over `pieces` = [(offset, n, need, done_index), ...] (floats, multiples of 4): piece i
        travels as soon as done[done_index] has reached `need` on this rank (need = 0: complete by stream order) and every
        other rank reports the same -- hbr_allreduce_peer_stream.  Enqueue it on a side stream AFTER the producer kernel
        (ops.hash_encode_bwd*_stream) has been enqueued on the compute stream: the two run side by side."""
        k = len(pieces)
        for off, n, _need, _idx in pieces:
            if off % 4 or n % 4 or off < 0 or n <= 0 or off + n > self.n:
                raise ValueError("pieces must be multiples of 4 floats inside the region")
        offs = (C.c_int64 * k)(*[int(p[0]) for p in pieces])
        ns = (C.c_int64 * k)(*[int(p[1]) for p in pieces])
        need = (C.c_uint * k)(*[int(p[2]) for p in pieces])
        idx = (C.c_int * k)(*[int(p[3]) for p in pieces])
        mc = C.c_void_p(self.multicast_ptr) if self.multicast_ptr else None
        dptr = C.c_void_p(done.data_ptr()) if done is not None else None
        check(lib().hbr_allreduce_peer_stream(self._bufs, self._flags, mc, self.rank, self.world, k, offs, ns, need, idx, dptr,
                                              float(scale), int(ctas), self._status_ptr, _lib.stream()))
        self.calls += 1

    def poll_status(self):
        """Enqueue a 4-byte copy of the status word to pinned host memory on the current stream (capturable; no sync)."""
        word = self._whole[self._status_off // 4: self._status_off // 4 + 1].view(torch.int32)
        self._status_host.copy_(word, non_blocking=True)

    def raise_if_failed(self):
        """Raises if a status copy that has ALREADY landed shows a failed exchange (a flag barrier timed out on some earlier
        all_reduce: this rank's gradients were left unreduced).  Never synchronises: detection lags by at most one step."""
        if int(self._status_host[0]) != 0:
            raise RuntimeError("peer-memory gradient all-reduce failed: a flag barrier timed out (a rank stalled or died); "
                               "gradients of that step were NOT reduced -- stop and restart the ranks")

    def timed_out(self) -> bool:
        """True if a flag barrier of any all_reduce so far gave up (synchronises)."""
        torch.cuda.synchronize()
        word = self._whole[self._status_off // 4: self._status_off // 4 + 1]
        return bool(word.view(torch.int32).item() != 0)

    def close(self):
        for p in self._imported:
            lib().hbr_peer_release(p)
        self._imported = []
        self.tensor = self._whole = None
        if self._own is not None:
            lib().hbr_peer_free(self._own)
            self._own = None
