"""Drop-in for the reference's helper.py: the free functions the trainers import with `from helper import *`.

Hot-path functions (`calc_color`, `hierarchical_sampling`) run on the sm_100a kernels in csrc/composite.cu;
setup helpers (`strat_sampler`, `get_od`, `find_bounding_box`) are a handful of torch ops evaluated with the
reference's exact expressions so RNG order and rounding match.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import ops
from .test_hash import MLP_3D  # noqa: F401  (helper.py:10 re-exports it through `from helper import *`)


# SDF mode on the dedicated kernels of csrc/sdf.cu (False: the composed tensor expressions, kept as the A/B for the tests)
SDF_KERNELS = True


class VarModel(nn.Module):
    """helper.py:13-21 (SDF mode's learnable sharpness)."""

    def __init__(self):
        super().__init__()
        self.b = nn.Parameter(torch.tensor(0.5))

    def forward(self, x):
        return 1 / (1 + torch.exp(-x * self.b))


def strat_sampler(tn, tf, num_samples: int, exp: Optional[bool] = False, device: Optional[str] = None) -> torch.Tensor:
    """helper.py:210-237.  ONE depth vector shared by all rays; RNG draw #1 = rand_like(t) on `device`."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if exp:
        t = torch.linspace(torch.log(tn), torch.log(tf), num_samples, device=device)
        t = t + (torch.rand_like(t) * (torch.log(tf) - torch.log(tn)) / num_samples)
        return torch.exp(t)
    t = _linspace(tn, tf, num_samples, device)
    u = torch.rand_like(t)                                                # RNG draw #1
    if t.is_cuda and t.dtype == torch.float32 and not torch.is_tensor(num_samples):
        span = tf - tn                                                    # CPU 0-dim tensors in the reference's callers
        if not (torch.is_tensor(span) and span.is_cuda):
            return ops.strat_depths(t, u, float(span), float(num_samples))   # same three roundings, one kernel
    return t + (u * (tf - tn) / num_samples)


_LIN = {}


def _linspace(tn, tf, num_samples, device):
    """torch.linspace(tn, tf, num_samples) -- a constant of the run for the reference's callers (near / far are fixed CPU
    scalars): computed once per (tn, tf, S, device) and reused, never written to."""
    if (torch.is_tensor(tn) and tn.is_cuda) or (torch.is_tensor(tf) and tf.is_cuda):
        return torch.linspace(tn, tf, num_samples, device=device)
    key = (float(tn), float(tf), int(num_samples), str(torch.device(device)), torch.cuda.current_device() if torch.cuda.is_available() else -1)
    lin = _LIN.get(key)
    if lin is None:
        lin = _LIN[key] = torch.linspace(tn, tf, num_samples, device=device)
    return lin


def hierarchical_sampling(rays_o, rays_d, z_vals, weights, n_samples: int, tn, tf, perturb: bool = False,
                          device: str = "cuda", _u=None, _u_cand=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """helper.py:23-51.  Returns (rays (R,2S,3), combined depths (R,2S)).

    RNG draws #2 (rand(R,S)) and #3 (rand(S)) happen here in the reference's order; tests inject them through
    _u / _u_cand.  Like the reference, negative weights are zeroed IN PLACE in the caller's tensor (:36)."""
    w = weights.squeeze(-1) if weights.dim() == 3 else weights
    if not w.is_cuda:
        raise RuntimeError("hierarchical_sampling needs CUDA tensors (there is no CPU fallback)")
    dev = w.device
    u = torch.rand(w.shape, device=dev) if _u is None else _u
    u_c = torch.rand(n_samples, device=dev) if _u_cand is None else _u_cand
    cand = u_c * (tf - tn) + tn                                          # :43, one shared unsorted candidate vector
    if w.dtype == torch.float32 and w.is_contiguous():
        combined = ops.hier_sample(w, z_vals, u, cand, clamp_in_place=True)
    else:
        w32 = w.float().contiguous()
        combined = ops.hier_sample(w32, z_vals, u, cand, clamp_in_place=True)
        with torch.no_grad():
            w.copy_(w32)
    rays = ops.ray_points(rays_o, rays_d, combined)                      # :48
    return rays, combined


def calc_color(t, rgb, sigma, dir_norm, use_sdf: bool = False, var_model=None, rays=None, model=None, encoder=None,
               device: str = "cuda"):
    """helper.py:53-107.  NeRF mode returns (Cr (R,3), wts (R,S,1), None); SDF mode (use_sdf) also the eikonal norms.  fp16 inputs (autocast) are promoted
    to fp32 like the reference's mixed-dtype arithmetic.  The -10 clamp (:76) is applied inside the kernel; the
    caller's sigma tensor is left untouched."""
    if not sigma.is_cuda:
        raise RuntimeError("calc_color needs CUDA tensors (there is no CPU fallback)")
    if use_sdf:
        return _calc_color_sdf(t, rgb, sigma, var_model, rays, model, encoder)
    Cr, w = ops.CompositeSplit.apply(rgb, sigma, t, dir_norm, None)
    return Cr, w[..., None], None


def _calc_color_sdf(t, rgb, sigma, var_model, rays, model, encoder):
    """helper.py:76-89,102-107, SDF mode: alpha_i = relu(1 - phi(s_{i+1}) / phi(s_i)) with phi = var_model (a sigmoid of
    learnable sharpness), transmittance = exclusive cumprod(1 - alpha), plus the eikonal term |grad sdf| from central
    differences at the sample positions.  With the reference's own VarModel and the native MLP_3D / HashEncoder this is
    three launches each way -- hbr_composite_sdf_fwd/bwd (the -10 clamp of :76 applied inside, the caller's tensor left
    untouched like the NeRF branch) and the 6-point stencil of MLP_3D.eikonal_norms around ONE encoder + density-head
    pass; any other var_model / model / encoder takes the reference's tensor expressions on the device.  Like the reference
    it needs `rays` (so hierarchical=True, whose fine pass passes none, fails the same way: AttributeError on None.device)."""
    mlp = getattr(model, "module", model)                                # the reference reaches through nn.DataParallel (:87)
    if SDF_KERNELS and type(var_model) is VarModel and sigma.dim() == 2 and sigma.shape[1] <= 1024:
        Cr, w = ops.CompositeSdf.apply(rgb, sigma, var_model.b, False)
        if rays is None:                                                 # vol_renderer.py:242 passes none for the fine pass:
            raise AttributeError("'NoneType' object has no attribute 'device'")   # the reference dies at test_hash.py:90
        if isinstance(mlp, MLP_3D) and mlp._native and hasattr(encoder, "_flat_table"):
            norm = mlp.eikonal_norms(rays, encoder=encoder)
        else:
            norm = eikonal_value(mlp.finite_difference_normals_approximator(rays, encoder=encoder))
        return Cr, w[..., None], norm
    sigma[sigma < -10] = -10                                             # :76 (in place, as the reference)
    phi = var_model(sigma)
    alpha = torch.zeros_like(sigma)
    alpha[..., :-1] = 1 - phi[..., 1:] / phi[..., :-1]
    alpha = torch.nn.functional.relu(alpha)
    T = cumprod_exclusive(1 - alpha)
    grads = mlp.finite_difference_normals_approximator(rays, encoder=encoder)
    norm = eikonal_value(grads)
    wts = T[:, :, None] * alpha[:, :, None]
    Cr = torch.sum(T[:, :, None] * alpha[:, :, None] * rgb, dim=-2)
    return Cr, wts, norm


def get_od(H, W, K, c2w: torch.Tensor, find_inv: Optional[bool] = False):
    """helper.py:176-208: per-pixel ray origins, unit directions and the direction norms (>= 1).
    K is the INTEGER intrinsics matrix of train_hash2.py:67-72 (focal / principal point truncated)."""
    device = c2w.device
    if c2w.is_cuda and not find_inv and c2w.dim() == 3:
        # one kernel instead of meshgrid + bmm + norm + div: rays of every pixel of every view, in the reference's order
        B, hw = c2w.shape[0], int(H) * int(W)
        o, d, n, _ = ops.ray_gen(c2w, int(H), int(W), K)
        return o.view(B, hw, 3), d.view(B, hw, 3), n.view(B, hw, 1)
    i, j = torch.meshgrid(torch.arange(W, device=device), torch.arange(H, device=device), indexing="xy")
    i = ((i - K[0, 2]) / K[0, 0]).reshape(-1)
    j = ((j - K[1, 2]) / K[1, 1]).reshape(-1)
    dirs = torch.stack((i, -j, -torch.ones_like(i)), axis=-1)
    if find_inv:
        rays_d = (torch.linalg.inv(c2w[..., :3, :3]) @ dirs.mT).mT
    else:
        rays_d = (c2w[..., :3, :3] @ dirs.mT).mT
    rays_o = (c2w[..., :3, 3:4].mT).expand(-1, rays_d.shape[1], -1)
    n = torch.norm(rays_d, dim=-1, keepdim=True)
    return rays_o, rays_d / n, n


def _bbox_update(rays, min_bound, max_bound):
    mn = rays.min(dim=0).values
    mx = rays.max(dim=0).values
    return torch.minimum(min_bound, mn.to(min_bound.dtype)), torch.maximum(max_bound, mx.to(max_bound.dtype))


def find_bounding_box(data_loader, near, far, K, num_samples=64, exp=False, device=None):
    """helper.py:109-141: bbox of the ray end points at depths {near, far+1.5} over all cameras.
    Returns (max_bound, min_bound)."""
    if device is None:
        device = K.device
    W = 2 * K[0, 2]
    H = 2 * K[1, 2]
    if exp:
        t = torch.from_numpy(np.asarray([near, far * torch.exp(torch.as_tensor(torch.log(far) - torch.log(near)) / num_samples)])).to(device)
    else:
        t = torch.from_numpy(np.asarray([near, far + 1.5])).to(device)
    min_bound = torch.ones(3, device=device) * 1e7
    max_bound = torch.ones(3, device=device) * (-1e7)
    if torch.device(device).type == "cuda" and not exp:
        # device path: one reduction kernel per loader batch, nothing materialised (hbr_ray_bbox)
        bounds = None
        for batch in data_loader:
            _, c2w, _ = batch
            bounds = ops.ray_bbox(c2w.to(device), int(H), int(W), K, float(t[0]), float(t[1]), bounds)
        return bounds[3:].clone(), bounds[:3].clone()
    with torch.no_grad():
        for batch in data_loader:
            _, c2w, _ = batch
            c2w = c2w.to(device)
            rays_o, rays_d, _ = get_od(H, W, K, c2w)
            rays = (rays_o[..., None, :] + rays_d[..., None, :] * t[None, :, None]).reshape(-1, 3)
            min_bound, max_bound = _bbox_update(rays, min_bound, max_bound)
    return max_bound, min_bound


def find_bounding_box2(data_loader, near, far, K, num_samples=64, exp=False, device=None):
    """helper.py:143-174: same, for loaders that already yield (rays_o, rays_d, _, _)."""
    if device is None:
        device = K.device
    t = torch.from_numpy(np.asarray([near, far + 1.5])).to(device)
    min_bound = torch.ones(3, device=device) * 1e7
    max_bound = torch.ones(3, device=device) * (-1e7)
    with torch.no_grad():
        for rays_o, rays_d, _, _ in data_loader:
            rays = (rays_o[..., None, :] + rays_d[..., None, :] * t[None, :, None]).reshape(-1, 3)
            min_bound, max_bound = _bbox_update(rays.to(device), min_bound, max_bound)
    return max_bound, min_bound


def cumprod_exclusive(tensor: torch.Tensor) -> torch.Tensor:
    """helper.py:268-291."""
    c = torch.roll(torch.cumprod(tensor, -1), 1, -1)
    c[..., 0] = 1.0
    return c


def eikonal_value(grads):
    return torch.sqrt(torch.sum(grads ** 2, dim=-1))


def eikonal_loss(norm):
    return torch.mean((norm - 1) ** 2)


def calc_psnr(pred, target):
    mse = torch.mean((pred - target) ** 2)
    return 10 * torch.log10(1.0 / mse)


def plot_grad_flow(named_parameters):
    """helper.py:239-266 (debug plot; needs matplotlib, which is optional here)."""
    import matplotlib.pyplot as plt  # noqa
    ave, mx, layers = [], [], []
    for n, p in named_parameters:
        if p.requires_grad and "bias" not in n:
            layers.append(n)
            ave.append(p.grad.abs().mean().cpu())
            mx.append(p.grad.abs().max().cpu())
    plt.bar(np.arange(len(mx)), mx, alpha=0.1, lw=1, color="c")
    plt.bar(np.arange(len(mx)), ave, alpha=0.1, lw=1, color="b")
    plt.xticks(range(0, len(ave), 1), layers, rotation="vertical")
    plt.show()
