"""Drop-in for the reference's encoder.py: `PositionalEncoder` (encoder.py:8-33) on the sm_100a kernel
csrc/composite.cu:dir_encode_kernel.  Output (..., d_model, 2*num_freq): [sin(2 x k)]_k || [cos(2 x k)]_k with
LINEAR k = 0..num_freq-1, exactly the layout the callers reshape to (..., d_model*2*num_freq)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _DirEncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, num_freq):
        ctx.save_for_backward(x2d)
        ctx.num_freq = num_freq
        return ops.dir_encode(x2d, num_freq)

    @staticmethod
    def backward(ctx, g):
        # d/dx sin(2xk) = 2k cos(2xk), d/dx cos(2xk) = -2k sin(2xk); rarely needed (directions are data)
        (x,) = ctx.saved_tensors
        nf = ctx.num_freq
        k = torch.arange(nf, device=x.device, dtype=torch.float32)
        a = 2 * x.float().unsqueeze(-1) * k
        g = g.reshape(x.shape[0], x.shape[1], 2 * nf)
        gx = (g[..., :nf] * (2 * k) * torch.cos(a) - g[..., nf:] * (2 * k) * torch.sin(a)).sum(-1)
        return gx.to(x.dtype), None


class PositionalEncoder(nn.Module):
    def __init__(self, d_model, num_freq=10):
        super().__init__()
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.d_model = d_model
        self.max_seq_len = num_freq
        self.sinus_in = torch.arange(0, self.max_seq_len, dtype=torch.int8).to(self.device)[None, None, :]   # encoder.py:16-17

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("PositionalEncoder.forward needs CUDA tensors (there is no CPU fallback)")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        in_dtype = x2.dtype
        if in_dtype not in (torch.float32, torch.float16):
            x2 = x2.float()
        out = _DirEncodeFn.apply(x2, self.max_seq_len) if x2.requires_grad else ops.dir_encode(x2, self.max_seq_len)
        if in_dtype == torch.float16:
            out = out.half()                                       # the reference returns the input dtype
        return out.reshape(lead + (x.shape[-1], 2 * self.max_seq_len)).reshape(lead + (-1,))
