"""CUDA-graph capture of one training step of the hot path.

A train_hash2.py step at 4096 rays is ~35 kernel launches of 2-250 us each; driven eagerly from Python the host
needs ~1 ms to enqueue them, more than the GPU needs to run them.  `GraphedStep` captures

    vol_render(...)  ->  loss_fn(Cr, Cf, gt)  ->  loss.backward()

once (fixed ray count / sample count) and replays it: one `cudaGraphLaunch` per step.  Inputs are copied into
static device buffers before each replay; the loss and the parameter gradients live in static buffers that every
replay overwrites (the parameters' `.grad` tensors are allocated inside the graph's memory pool during capture, the
documented whole-network-capture pattern of torch.cuda.graphs).  torch's CUDA generator is graph-safe, so the
reference's RNG draws (strat_sampler, hierarchical_sampling) still advance on every replay.
An optimiser may be stepped eagerly after the replay (its inputs, the `.grad` tensors, are static), or -- with
`optimizers=[optim.FusedAdam(..., capturable=True), ...]` -- inside the captured graph, so that one replay is a complete
train_hash2.py iteration (:218-239) including both optimiser steps.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch


def default_loss(Cr, Cf, gt):
    """train_hash2.py:221 with the MSE criterion of :177.  Without hierarchical sampling vol_render returns Cf = Cr (the
    same tensor, vol_renderer.py:244): mse + mse of one tensor is exactly 2 * mse (x + x is exact in binary floating
    point), which halves the loss kernels."""
    from . import ops
    if Cr.is_cuda and Cr.dtype == torch.float32 and gt.dtype == torch.float32 and Cr.shape == gt.shape:
        # one kernel per direction (ops.MsePair) instead of ~10 elementwise / reduce launches
        return ops.mse_pair(Cr, None, gt, 2.0) if Cf is Cr else ops.mse_pair(Cr, Cf, gt)
    if Cf is Cr:
        return 2.0 * torch.nn.functional.mse_loss(Cr, gt)
    return torch.nn.functional.mse_loss(Cr, gt) + torch.nn.functional.mse_loss(Cf, gt)


class GraphedStep:
    def __init__(self, renderer, model, params: Iterable[torch.nn.Parameter], n_rays: int, num_samples: int,
                 hierarchical: bool, device, loss_fn: Callable = default_loss, autocast: bool = True, warmup: int = 3,
                 source=None, autocast_dtype: torch.dtype = torch.bfloat16, optimizers: Iterable = ()):
        """source: optional rays.DeviceRayDataset.  Its sampler (ray ids from the graph-safe device generator ->
        hbr_ray_gen) is then captured in front of the step, so a replay draws a fresh batch from the resident views by
        itself: call the object with no arguments; nothing crosses PCIe but the graph launch."""
        self.renderer, self.model = renderer, model
        self.source = source
        self.params = list(params)
        self.num_samples, self.hierarchical = int(num_samples), bool(hierarchical)
        self.loss_fn, self.autocast, self.autocast_dtype = loss_fn, autocast, autocast_dtype
        # optimisers stepped INSIDE the captured graph (train_hash2.py:226-239 in one replay): optim.FusedAdam(capturable=True)
        # keeps step count and learning rate on the device; note that the warm-up iterations of capture() are real steps
        # on whatever the static input buffers hold: load() a real batch before capture()
        self.optimizers = list(optimizers)
        for o in self.optimizers:
            if not getattr(o, "capturable", False):
                raise ValueError("optimisers captured in the step's graph must be capturable (optim.FusedAdam(capturable=True))")
        dev = torch.device(device)
        # one packed static buffer [rays_o (R,3) | rays_d (R,3) | dir_norm (R,1) | gt (R,3)]: a batch packed the same way
        # (pack_batch) is loaded with ONE copy instead of four
        self.n_rays = int(n_rays)
        self.packed = torch.zeros(10 * n_rays, device=dev)
        R = n_rays
        self.rays_o = self.packed[0:3 * R].view(R, 3)
        self.rays_d = self.packed[3 * R:6 * R].view(R, 3)
        self.dir_norm = self.packed[6 * R:7 * R].view(R, 1)
        self.gt = self.packed[7 * R:10 * R].view(R, 3)
        self.rays_d[:, 2] = 1.0
        self.dir_norm.fill_(1.0)
        self.loss: Optional[torch.Tensor] = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._warmup = warmup

    def _step(self):
        rays_o, rays_d, dir_norm, gt = self.rays_o, self.rays_d, self.dir_norm, self.gt
        if self.source is not None:
            rays_o, rays_d, dir_norm, gt = self.source.sample(self.n_rays)
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast):
            Cr, Cf, _ = self.renderer.vol_render(self.model, rays_d, rays_o, num_samples=self.num_samples,
                                                 update_mask=False, dir_norm=dir_norm, hierarchical=self.hierarchical)
            loss = self.loss_fn(Cr, Cf, gt)
        loss.backward()
        for o in self.optimizers:
            o.step()
        return loss

    @staticmethod
    def pack_batch(rays_o, rays_d, dir_norm, gt, pin: bool = False) -> torch.Tensor:
        """(rays_o, rays_d, dir_norm, gt) -> the flat (10 R,) layout of the static input buffer."""
        flat = torch.cat([rays_o.reshape(-1), rays_d.reshape(-1), dir_norm.reshape(-1), gt.reshape(-1)]).float().contiguous()
        return flat.pin_memory() if pin and not flat.is_cuda else flat

    def load(self, rays_o, rays_d=None, dir_norm=None, gt=None, non_blocking: bool = True):
        """Copy a batch (device or pinned-host tensors) into the static input buffers: either the four tensors or one
        packed tensor from pack_batch()."""
        if rays_d is None:
            self.packed.copy_(rays_o, non_blocking=non_blocking)
            return
        self.rays_o.copy_(rays_o, non_blocking=non_blocking)
        self.rays_d.copy_(rays_d, non_blocking=non_blocking)
        self.dir_norm.copy_(dir_norm.reshape(-1, 1), non_blocking=non_blocking)
        self.gt.copy_(gt, non_blocking=non_blocking)

    def capture(self):
        # the parameters' AccumulateGrad nodes may predate the capture stream (eager steps before capture): harmless
        # here (every step of the graph runs on the capture stream), so silence torch's per-backward warning about it
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self._warmup):
                for p in self.params:
                    p.grad = None
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        return self

    def __call__(self, rays_o=None, rays_d=None, dir_norm=None, gt=None) -> torch.Tensor:
        if self.graph is None:
            self.capture()
        if rays_o is not None:
            self.load(rays_o, rays_d, dir_norm, gt)
        self.graph.replay()
        return self.loss
