"""Device-resident training rays: replaces the data path of train_hash2.py:74-96,211-215.

The reference evaluates get_od for every pixel of every view on the CPU, concatenates (rays_o, rays_d, dir_norm, gt)
into a TensorDataset (2.56 GB for 100 views of 800x800), and feeds 4096-ray batches through a shuffled DataLoader with
8 worker processes and a host->device copy per step.  `DeviceRayDataset` keeps only the views (c2w + pixels: 192 MB as
uint8) in HBM and generates each batch where it is consumed with one kernel (hbr_ray_gen): no per-step host work and no
H2D.  Iterating it yields the batches of `DataLoader(TensorDataset(...), batch_size, shuffle=True)`: every ray once per
epoch in a fresh random order (torch.randperm on the device), the last batch partial.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch

from . import ops


class DeviceRayDataset:
    def __init__(self, images: torch.Tensor, c2w: torch.Tensor, K, H: Optional[int] = None, W: Optional[int] = None,
                 device=None, batch_size: int = 4096, shuffle: bool = True, generator: Optional[torch.Generator] = None):
        """images: (V,3,H,W) as the reference's datasets yield them (dataset.py:38-44, float in [0,1] after ToTensor) or
        (V,H,W,3); float32 or uint8 (uint8 is kept as uint8 and divided by 255 on the fly, as ToTensor would have).
        c2w: (V,4,4).  K: the 3x3 intrinsics the trainer builds (train_hash2.py:67-72)."""
        if device is None:
            device = "cuda"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceRayDataset needs a CUDA device (there is no CPU path)")
        if images.dim() != 4:
            raise ValueError("images must be (V,3,H,W) or (V,H,W,3)")
        if images.shape[1] == 3 and images.shape[-1] != 3:
            images = images.permute(0, 2, 3, 1)                       # train_hash2.py:81
        if images.dtype not in (torch.uint8, torch.float32):
            images = images.float()
        self.images = images.to(self.device).contiguous()
        V, h, w, _ = self.images.shape
        self.H, self.W = int(H or h), int(W or w)
        if (self.H, self.W) != (h, w):
            raise ValueError(f"images are {h}x{w}, H/W say {self.H}x{self.W}")
        self.c2w = c2w.to(self.device, torch.float32).contiguous()
        if self.c2w.shape != (V, 4, 4):
            raise ValueError("c2w must be (V,4,4) with one pose per image")
        self.K = torch.as_tensor(K).detach().cpu().clone()
        self.batch_size, self.shuffle, self.generator = int(batch_size), shuffle, generator

    @property
    def n_rays(self) -> int:
        return self.c2w.shape[0] * self.H * self.W

    def __len__(self) -> int:                                          # batches per epoch, like len(DataLoader)
        return -(-self.n_rays // self.batch_size)

    def rays(self, ray_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """(rays_o, rays_d, dir_norm, gt) of the given flat ray ids -- row ids of the reference's TensorDataset."""
        return ops.ray_gen(self.c2w, self.H, self.W, self.K, ray_ids=ray_ids.to(self.device), images=self.images)

    def sample(self, n: Optional[int] = None):
        """n rays drawn uniformly with replacement (no epoch bookkeeping; ids come from the device generator)."""
        n = self.batch_size if n is None else int(n)
        ids = torch.randint(0, self.n_rays, (n,), device=self.device, generator=self.generator)
        return self.rays(ids)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]]:
        n = self.n_rays
        order = torch.randperm(n, device=self.device, generator=self.generator) if self.shuffle else None
        for b0 in range(0, n, self.batch_size):
            b1 = min(n, b0 + self.batch_size)
            if order is None:
                yield ops.ray_gen(self.c2w, self.H, self.W, self.K, first=b0, n_rays=b1 - b0, images=self.images)
            else:
                yield self.rays(order[b0:b1])

    def bounding_box(self, near, far, views_per_call: int = 0):
        """find_bounding_box (helper.py:109-141) over all views: (max_bound, min_bound), t in {near, far + 1.5}.  The
        reference re-derives the image size from the principal point (W = 2*K[0,2], H = 2*K[1,2], :114-115)."""
        W, H = int(2 * self.K[0, 2]), int(2 * self.K[1, 2])
        b = ops.ray_bbox(self.c2w, H, W, self.K, float(near), float(far) + 1.5)
        return b[3:].clone(), b[:3].clone()
