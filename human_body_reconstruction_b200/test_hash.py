"""Drop-in for the model half of the reference's test_hash.py: `MLP_3D` (test_hash.py:20-105).

Same constructor, `forward(x, viewdirs=None, mask=None)`, state_dict keys (`sig_model.{0,2,4}.*`,
`col_model.{0,2,4}.*`) and RNG consumption at construction.  The twelve parameter tensors are views into
one flat fp32 buffer (state_dict order) that the kernels read directly.

Precision: fp32 CUDA-core kernels by default (nerf2mesh.py runs the field without autocast); under
`torch.autocast` the forward/backward run on the tcgen05 tensor-core kernels with bf16 operands and fp32
accumulation (the reference runs fp16 there, train_hash2.py:218).

Only the shape the reference instantiates -- num_sig=2, num_col=2, h_size=64 (train_hash2.py:127,
nerf2mesh.py:57) -- has a native implementation; anything else raises NotImplementedError.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import MlpDims


class _MlpFn(torch.autograd.Function):
    """feat (N,in0), dirs (N/dir_group, d_view) -> (N,4) [rgb, sigma]."""

    @staticmethod
    def forward(ctx, feat, dirs, dir_group, mlp, use_tc, *params):
        flat = mlp._flat_params()
        dims = mlp._dims()
        train = any(ctx.needs_input_grad)
        feat = feat.float().contiguous()
        dirs = dirs.float().contiguous()
        if use_tc:
            out, act = ops.mlp_fwd_tc(feat, dirs, dir_group, flat, dims, keep_act=train)
        else:
            out, act = ops.mlp_fwd_f32(feat, dirs, dir_group, flat, dims, keep_act=train)
        ctx.mlp, ctx.dims, ctx.dir_group, ctx.use_tc = mlp, dims, dir_group, use_tc
        ctx.save_for_backward(feat, dirs, out if use_tc else act)     # the tensor-core backward reads the saved output
        return out

    @staticmethod
    def backward(ctx, dout):
        feat, dirs, act = ctx.saved_tensors
        mlp = ctx.mlp
        flat = mlp._flat_params()
        dflat = mlp._grad_buffer.zero_() if mlp._grad_buffer is not None else torch.zeros_like(flat)
        dout = dout.float().contiguous()
        want_dfeat, want_ddirs = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if ctx.use_tc:
            dfeat, ddirs = ops.mlp_bwd_tc(feat, dirs, ctx.dir_group, flat, ctx.dims, act.detach(), dout, want_dfeat, want_ddirs,
                                          dflat)
        else:
            dfeat, ddirs = ops.mlp_bwd_f32(feat, dirs, ctx.dir_group, flat, ctx.dims, dout, act, want_dfeat, want_ddirs, dflat)
        mlp._publish_grad(dflat)
        return (dfeat, ddirs, None, None, None) + tuple(mlp._grad_views(dflat))


class MLP_3D(nn.Module):
    def __init__(self, num_sig=3, num_col=2, h_size=64, d_view=3, L=16, F=2, E=0, use_sdf=False, max_bound=1.0,
                 min_bound=-1.0):
        super().__init__()
        self.d_view = d_view
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        self.max_bound = torch.as_tensor(max_bound).to(dev)          # test_hash.py:25-26
        self.min_bound = torch.as_tensor(min_bound).to(dev)
        sig = [nn.Linear(L * F + E, h_size), nn.ReLU()]
        for i in range(num_sig):
            if i == num_sig - 1:
                sig.append(nn.Linear(h_size, 1 + 15))
            else:
                sig += [nn.Linear(h_size, h_size), nn.ReLU()]
        self.sigmoid = nn.Sigmoid()
        self.sig_model = nn.Sequential(*sig)
        self.relu, self.elu, self.lrelu = nn.ReLU(), nn.ELU(), nn.LeakyReLU()
        col = [nn.Linear(15 + d_view, h_size), nn.ReLU()]
        for i in range(num_col):
            if i == num_col - 1:
                col.append(nn.Linear(h_size, 3))
            else:
                col += [nn.Linear(h_size, h_size), nn.ReLU()]
        self.col_model = nn.Sequential(*col)
        self.use_sdf = use_sdf
        self._native = (num_sig == 2 and num_col == 2 and h_size == 64 and L * F + E <= 64 and 15 + d_view <= 64)
        self._in0 = L * F + E
        self._grad_hooks = []
        self._grad_buffer = None     # persistent (peer-mapped) flat gradient buffer, see dist.PeerGradAllReduce
        self._flat = None
        if self._native:
            self._reflatten()

    # -- flat parameter storage ----------------------------------------------------------------------------
    def _ordered(self):
        ps = []
        for seq in (self.sig_model, self.col_model):
            for m in seq:
                if isinstance(m, nn.Linear):
                    ps += [m.weight, m.bias]
        return ps

    def _reflatten(self):
        ps = self._ordered()
        n = sum(p.numel() for p in ps)
        flat = torch.empty(n, device=ps[0].device, dtype=torch.float32)
        o = 0
        with torch.no_grad():
            for p in ps:
                v = flat[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                o += p.numel()
        self._flat = flat

    def _flat_params(self) -> torch.Tensor:
        flat = self._flat
        ok = flat is not None
        if ok:
            o = flat.data_ptr()
            for p in self._ordered():
                if p.data_ptr() != o or p.device != flat.device or p.dtype != torch.float32:
                    ok = False
                    break
                o += p.numel() * 4
        if not ok:
            self._reflatten()
        return self._flat

    def _grad_views(self, dflat):
        o, out = 0, []
        for p in self._ordered():
            out.append(dflat[o:o + p.numel()].view(p.shape))
            o += p.numel()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flat = None
        return out

    def _dims(self) -> MlpDims:
        return MlpDims(self._in0, self.d_view)

    def _publish_grad(self, g):
        for h in self._grad_hooks:
            h(g)

    def _check_native(self):
        if not self._native:
            raise NotImplementedError("native MLP_3D supports num_sig=2, num_col=2, h_size=64 (the reference's configuration)")
        if self.use_sdf:
            raise NotImplementedError("use_sdf=True (SDF mode, SURVEY 8f row 4) is not implemented on the CUDA path yet")

    # -- evaluation entry used by Volume_Renderer (one direction row per ray) ---------------------------------
    def field(self, feat, dirs_enc, dir_group, use_tc=None):
        self._check_native()
        if use_tc is None:
            use_tc = torch.is_autocast_enabled()
        use_tc = bool(use_tc) and ops.HAS_TC
        return _MlpFn.apply(feat, dirs_enc, int(dir_group), self, bool(use_tc), *self._ordered())

    def forward(self, x, viewdirs=None, mask=None):
        self._check_native()
        if not x.is_cuda:
            raise RuntimeError("MLP_3D.forward needs CUDA tensors (there is no CPU fallback)")
        if viewdirs is None:
            if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
                # the density-only branch (test_hash.py:73-77) is inference-only in the reference's callers
                z = torch.zeros((x.shape[0], self.d_view), device=x.device)
                density = self.field(x, z, 1, use_tc=False)[:, 3:4]
            else:
                density, _ = ops.mlp_fwd_f32(x, None, 1, self._flat_params(), self._dims(), keep_act=False)
            return density * mask if mask is not None else density
        out = self.field(x, viewdirs, 1)
        if mask is not None:
            out = out * mask[..., None]
        return out
