"""Fused Adam / AdamW for the flat parameter buffers.

train_hash2.py:141-142 builds `Adam(encoder.Embedding_list.parameters(), lr=.05)` and `AdamW(nerf.parameters(), lr=.005)`
and steps both every iteration (:227-228); on the 16.8 M-entry table torch's multi-tensor Adam is seven elementwise passes
(470 MB of HBM traffic dominated by re-reads).  `FusedAdam` has the same constructor arguments, state_dict layout
(`exp_avg`, `exp_avg_sq`, `step` per parameter) and arithmetic, but updates every group with ONE kernel pass per
contiguous buffer: parameters that are adjacent slices of one allocation -- the encoder's level tables, the MLP's twelve
tensors -- and whose gradients are adjacent too (this package's backward hands autograd views of one flat buffer) are
merged into a single launch.  No CPU path: CUDA fp32 parameters only.
"""
from __future__ import annotations

import ctypes as C
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
    def __init__(self, params: Iterable, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled)
        super().__init__(params, defaults)
        self._flat_state = {}                    # (group index, run start) -> (exp_avg, exp_avg_sq) flat buffers
        self._run_cache = {}                     # group index -> (parameter offsets, runs) of the previous step

    @torch.no_grad()
    def step(self, closure=None, grad_scaler=None, inv_scale: float = 1.0, found_inf: torch.Tensor | None = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                        and p.grad.dtype == torch.float32):
                    raise RuntimeError("FusedAdam updates contiguous fp32 CUDA parameters only (there is no CPU fallback)")
            b1, b2 = group["betas"]
            for start, count, total in self._group_runs(gi, ps):
                key = (gi, ps[start].data_ptr(), total)
                st = self._flat_state.get(key)
                if st is None:
                    m = torch.zeros(total, device=ps[start].device, dtype=torch.float32)
                    v = torch.zeros_like(m)
                    o = 0
                    for p in ps[start:start + count]:       # per-parameter views: the torch.optim state_dict layout
                        s = self.state[p]
                        if "exp_avg" in s:                  # state loaded from a checkpoint
                            m[o:o + p.numel()].copy_(s["exp_avg"].reshape(-1))
                            v[o:o + p.numel()].copy_(s["exp_avg_sq"].reshape(-1))
                        s["exp_avg"] = m[o:o + p.numel()].view_as(p)
                        s["exp_avg_sq"] = v[o:o + p.numel()].view_as(p)
                        s.setdefault("step", 0)
                        o += p.numel()
                    st = self._flat_state[key] = (m, v)
                m, v = st
                step_no = int(self.state[ps[start]]["step"]) + 1
                for p in ps[start:start + count]:
                    self.state[p]["step"] = step_no
                with torch.cuda.device(ps[start].device):
                    check(lib().hbr_adam_step(ptr(ps[start]), ptr(ps[start].grad), ptr(m), ptr(v), total, float(group["lr"]),
                                              float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                              1 if group["decoupled"] else 0, step_no, float(inv_scale), ptr(found_inf),
                                              stream()))
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
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=True)
