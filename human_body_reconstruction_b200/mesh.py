"""The nerf2mesh.py hot loop on the GPU: density-grid query (nerf2mesh.py:27-40,69-87) and marching cubes
(nerf2mesh.py:98, third-party torchmcubes in the reference).

`density_grid` never materialises the res^3 x 3 position tensor on the host: positions are generated on the
device per chunk with numpy-1.23 linspace semantics (float64, end point pinned) and rounded to fp16 exactly
like `grid.to(torch.float16)`.  Slab sharding across ranks = one contiguous range of the flat grid per rank
(axis 0 of density[i,j,k], i.e. the y coordinate under meshgrid's 'xy' indexing).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .encoder import PositionalEncoder
from .hash_encoding import HashEncoder
from .test_hash import MLP_3D


def _unwrap(m):
    return m.module if isinstance(m, nn.DataParallel) else m


def view_dir_encoding(dir_encoder: PositionalEncoder, device) -> torch.Tensor:
    """PositionalEncoder((0,0,1) in fp16) as fp32 values (nerf2mesh.py:69-70,81): one row for every grid point."""
    v = torch.zeros((1, 3), device=device, dtype=torch.float16)
    v[..., 2] = 1.0
    return ops.dir_encode(v, dir_encoder.max_seq_len)


def density_grid(encoder: HashEncoder, nerf, dir_encoder: Optional[PositionalEncoder], min_bound, max_bound, res: int,
                 i_begin: int = 0, i_end: Optional[int] = None, chunk: int = 1 << 19, cuda_core_mlp: bool = False) -> torch.Tensor:
    """Field values on planes [i_begin, i_end) of the res^3 grid.

    With a dir_encoder: (planes, res, res, 4) = [rgb, density], the layout nerf2mesh.py:86-87 saves to
    density_grid_w_rgb.npy.  With dir_encoder=None: (planes, res, res) density only (all marching cubes needs,
    nerf2mesh.py:95)."""
    mlp = _unwrap(nerf)
    if not isinstance(encoder, HashEncoder) or not isinstance(mlp, MLP_3D):
        raise TypeError("density_grid needs the native HashEncoder and MLP_3D")
    mlp._check_native()
    i_end = res if i_end is None else i_end
    table = encoder._flat_table()
    if not table.is_cuda:
        raise RuntimeError("density_grid needs the encoder on a CUDA device (there is no CPU fallback)")
    p0, count = i_begin * res * res, (i_end - i_begin) * res * res
    denc = view_dir_encoding(dir_encoder, table.device) if dir_encoder is not None else None
    with torch.no_grad():
        out = ops.grid_density([float(v) for v in min_bound], [float(v) for v in max_bound], res, p0, count, table,
                               encoder._geom(), mlp._flat_params(), mlp._dims(), denc, chunk=chunk, cuda_core_mlp=cuda_core_mlp)
    return out.view((i_end - i_begin, res, res, 4) if denc is not None else (i_end - i_begin, res, res))


def marching_cubes_counts(density: torch.Tensor, iso: float, i_begin: int = 0, i_end: Optional[int] = None) -> Tuple[int, int]:
    """(vertices, triangles) owned by planes [i_begin, i_end): welded vertices = iso-crossing grid edges."""
    c = ops.mc_count(density, iso, i_begin, i_end).tolist()
    return int(c[0]), int(c[1])


def marching_cubes(density: torch.Tensor, iso: float):
    """torchmcubes.marching_cubes-shaped call (nerf2mesh.py:98): (verts (V,3) fp32 in grid-index coordinates along
    axes (0,1,2), faces (F,3) int32).  Inside test: density < iso.  One welded vertex per crossing edge."""
    nv, nt = marching_cubes_counts(density, iso)
    verts, faces, _ = ops.mc_emit(density, iso, nv, nt)
    return verts, faces


def _device_for(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("marching cubes / grid_interp run on a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def marching_cubes_xyz(vol: torch.Tensor, iso: float):
    """torchmcubes.marching_cubes as nerf2mesh.py:95-98 calls it: `vol` may live on the CPU (it does there) -- it is
    uploaded, meshed on the GPU and the result returned on vol's device.  Vertices are (x, y, z) = (index along axis 2,
    axis 1, axis 0), torchmcubes' order for a volume indexed [z, y, x]; faces (F,3) int32."""
    dev = _device_for(vol)
    d = vol.detach().to(dev, torch.float32).contiguous()
    with torch.cuda.device(dev):
        verts, faces = marching_cubes(d, iso)
        verts = verts.flip(-1).contiguous()
    return verts.to(vol.device), faces.to(vol.device)


def grid_interp(vol: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """torchmcubes.grid_interp (nerf2mesh.py:99): vol (C,Nz,Ny,Nx) or (Nz,Ny,Nx), points (V,3) as (x,y,z) -> (V,C)."""
    dev = _device_for(vol if vol.is_cuda else points)
    v = vol.detach().to(dev, torch.float32)
    if v.dim() == 3:
        v = v[None]
    with torch.cuda.device(dev):
        out = ops.grid_interp(v.contiguous(), points.detach().to(dev, torch.float32).contiguous())
    return out.to(points.device)
