"""Fused Adam / AdamW for the flat parameter buffers.

train_hash2.py:141-142 builds `Adam(encoder.Embedding_list.parameters(), lr=.05)` and `AdamW(nerf.parameters(), lr=.005)`
and steps both every iteration (:227-228); on the 16.8 M-entry table torch's multi-tensor Adam is seven elementwise passes
(470 MB of HBM traffic dominated by re-reads).  `FusedAdam` has the same constructor arguments, state_dict layout
(`exp_avg`, `exp_avg_sq`, `step` per parameter) and arithmetic, but updates every group with ONE kernel pass per
contiguous buffer: parameters that are adjacent slices of one allocation -- the encoder's level tables, the MLP's twelve
tensors -- and whose gradients are adjacent too (this package's backward hands autograd views of one flat buffer) are
merged into a single launch.  No CPU path: CUDA fp32 parameters only.

Two modes:
* default: step count and learning rate are host numbers (as torch.optim.Adam's default path);
* `capturable=True`, or whenever a GradScaler hands over `grad_scale` / `found_inf` tensors (the optimiser declares
  `_step_supports_amp_scaling`, so `scaler.step(opt)` calls `step()` directly and never synchronises): step count, learning
  rate, gradient scale and the skip flag all live on the device (hbr_adam_tick + hbr_adam_step_dev).  The step can then be
  captured inside the CUDA graph of the training step (graph.GraphedStep(optimizers=...)); a schedule changes the learning
  rate of a captured optimiser through `set_lr()` / `sync_lr()` (one tiny fill, outside the graph); a skipped step does
  not advance the bias correction.
"""
from __future__ import annotations

from typing import Iterable, List

import torch

from ._lib import check, lib, ptr, stream


def _runs(params: List[torch.Tensor]):
    """Maximal runs of parameters whose data AND grads are back-to-back in memory -> [(first_index, count, numel)]."""
    runs, i = [], 0
    while i < len(params):
        j, total = i, params[i].numel()
        while (j + 1 < len(params)
               and params[j + 1].data_ptr() == params[j].data_ptr() + params[j].numel() * 4
               and params[j + 1].grad.data_ptr() == params[j].grad.data_ptr() + params[j].numel() * 4):
            j += 1
            total += params[j].numel()
        runs.append((i, j - i + 1, total))
        i = j + 1
    return runs


class FusedAdam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True            # GradScaler.step() hands over grad_scale / found_inf instead of syncing

    def __init__(self, params: Iterable, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False,
                 capturable: bool = False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled)
        super().__init__(params, defaults)
        self.capturable = bool(capturable)
        self._flat_state = {}                    # (group index, run start ptr, numel) -> (exp_avg, exp_avg_sq) flat buffers
        self._run_cache = {}                     # group index -> (parameter offsets, runs) of the previous step
        self._dev = {}                           # device -> {"step": int64[1], "lr": {group index: float32[1]}}

    # -- state handling -------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """The flat moment buffers are rebuilt from the loaded per-parameter state at the next step (a cache that outlived
        a load would silently keep the old moments)."""
        super().load_state_dict(state_dict)
        self._flat_state.clear()
        self._run_cache.clear()
        self._dev.clear()

    def _device_scalars(self, device):
        d = self._dev.get(device)
        if d is None:
            d = self._dev[device] = {"step": torch.zeros(1, dtype=torch.int64, device=device), "lr": {}}
        return d

    def set_lr(self, lr: float, group: int | None = None):
        """Learning rate of a group (all groups by default): the host value a scheduler sees AND the device scalar a
        captured step reads."""
        for gi, g in enumerate(self.param_groups):
            if group is None or gi == group:
                g["lr"] = float(lr)
        self.sync_lr()

    def sync_lr(self):
        """Push param_groups[*]['lr'] (e.g. after scheduler.step()) to the device scalars of the capturable path."""
        for d in self._dev.values():
            for gi, t in d["lr"].items():
                t.fill_(float(self.param_groups[gi]["lr"]))

    def _flat_moments(self, gi, ps, start, count, total):
        """(exp_avg, exp_avg_sq) flat buffers of a run, with self.state[p] holding views of them.  A cached pair is reused
        only while every parameter's state still aliases it at the right offset (state replaced by load_state_dict, or a
        run that re-formed differently, is re-flattened from the per-parameter tensors -- never dropped)."""
        key = (gi, ps[start].data_ptr(), total)
        st = self._flat_state.get(key)
        if st is not None:
            m, v = st
            o, ok = 0, True
            for p in ps[start:start + count]:
                s = self.state.get(p, {})
                if ("exp_avg" not in s or s["exp_avg"].data_ptr() != m.data_ptr() + 4 * o
                        or s["exp_avg_sq"].data_ptr() != v.data_ptr() + 4 * o):
                    ok = False
                    break
                o += p.numel()
            if ok:
                return m, v
        m = torch.zeros(total, device=ps[start].device, dtype=torch.float32)
        v = torch.zeros_like(m)
        o = 0
        for p in ps[start:start + count]:       # per-parameter views: the torch.optim state_dict layout
            s = self.state[p]
            if "exp_avg" in s:                  # state loaded from a checkpoint, or left by another run layout
                m[o:o + p.numel()].copy_(s["exp_avg"].reshape(-1))
                v[o:o + p.numel()].copy_(s["exp_avg_sq"].reshape(-1))
            s["exp_avg"] = m[o:o + p.numel()].view_as(p)
            s["exp_avg_sq"] = v[o:o + p.numel()].view_as(p)
            s.setdefault("step", 0)
            o += p.numel()
        self._flat_state[key] = (m, v)
        return m, v

    # -- the step -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, inv_scale: float = 1.0, found_inf: torch.Tensor | None = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)              # set by GradScaler.step (tensor) for this call
        if found_inf is None:
            found_inf = getattr(self, "found_inf", None)
        on_device = self.capturable or grad_scale is not None or found_inf is not None
        ticked = set()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                        and p.grad.dtype == torch.float32):
                    raise RuntimeError("FusedAdam updates contiguous fp32 CUDA parameters only (there is no CPU fallback)")
            b1, b2 = group["betas"]
            dev = ps[0].device
            if on_device:
                d = self._device_scalars(dev)
                if dev not in ticked:                               # one tick per optimiser step and device
                    ticked.add(dev)
                    if not d.get("seeded"):                         # continue from a host-side count (mode switch / checkpoint)
                        host = max([int(self.state[p].get("step", 0)) if not torch.is_tensor(self.state[p].get("step", 0))
                                    else int(self.state[p]["step"].item()) for p in ps if p in self.state] or [0])
                        d["step"].fill_(host)
                        d["seeded"] = True
                    with torch.cuda.device(dev):
                        check(lib().hbr_adam_tick(ptr(d["step"]), ptr(found_inf), stream()))
                lr_dev = d["lr"].get(gi)
                if lr_dev is None:
                    lr_dev = d["lr"][gi] = torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev)
                elif not torch.cuda.is_current_stream_capturing():
                    lr_dev.fill_(float(group["lr"]))                # eager calls follow the scheduler by themselves
            for start, count, total in self._group_runs(gi, ps):
                m, v = self._flat_moments(gi, ps, start, count, total)
                with torch.cuda.device(dev):
                    if on_device:
                        for p in ps[start:start + count]:
                            self.state[p]["step"] = d["step"]
                        check(lib().hbr_adam_step_dev(ptr(ps[start]), ptr(ps[start].grad), ptr(m), ptr(v), total, ptr(lr_dev),
                                                      float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                                      1 if group["decoupled"] else 0, ptr(d["step"]), float(inv_scale),
                                                      ptr(grad_scale), ptr(found_inf), stream()))
                    else:
                        s0 = self.state[ps[start]]["step"]
                        step_no = (int(s0.item()) if torch.is_tensor(s0) else int(s0)) + 1
                        for p in ps[start:start + count]:
                            self.state[p]["step"] = step_no
                        check(lib().hbr_adam_step(ptr(ps[start]), ptr(ps[start].grad), ptr(m), ptr(v), total, float(group["lr"]),
                                                  float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                                  1 if group["decoupled"] else 0, step_no, float(inv_scale), ptr(found_inf),
                                                  stream()))
                # the kernels write the parameters through raw pointers: tell autograd's version counter (and anything
                # keyed on it, e.g. the MLP's cached operand image) that they changed
                torch.autograd.graph.increment_version(ps[start])
        return loss

    def _group_runs(self, gi, ps):
        """_runs(ps), cached: the layout of the previous step is reused after verifying that parameters and gradients
        still sit at the same offsets from their run heads (gradient buffers are fresh every backward, their adjacency
        normally is not)."""
        cached = self._run_cache.get(gi)
        if cached is not None and len(cached[0]) == len(ps):
            offs, runs = cached
            ok = True
            for start, count, _ in runs:
                p0, g0 = ps[start].data_ptr(), ps[start].grad.data_ptr()
                for j in range(start, start + count):
                    if ps[j].data_ptr() - p0 != offs[j] or ps[j].grad.data_ptr() - g0 != offs[j]:
                        ok = False
                        break
                if not ok:
                    break
            if ok:
                return runs
        runs = _runs(ps)
        offs = [0] * len(ps)
        for start, count, _ in runs:
            o = 0
            for j in range(start, start + count):
                offs[j] = o
                o += ps[j].numel() * 4
        self._run_cache[gi] = (offs, runs)
        return runs


class FusedAdamW(FusedAdam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, capturable: bool = False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=True, capturable=capturable)
