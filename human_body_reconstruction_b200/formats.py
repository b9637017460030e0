"""On-disk formats either side of the hot path, so data and checkpoints move freely between the
reference and this package.  Plain host-side I/O -- nothing here is on the GPU path.

* scene description: Blender-synthetic `transforms_*.json` (dataset.py:9-44: `camera_angle_x`, frames with `file_path`
  (no extension), `transform_matrix`, `rotation`) and the instant-ngp/nerfstudio flavour of the human captures
  (dataset_new.py:9-46: `w, h, fl_x, fl_y, cx, cy`, frames with `file_path` (with extension), `transform_matrix`,
  `sharpness`)  ->  `read_transforms` / `load_scene`
* intrinsics: the trainer's 3x3 int64 K (train_hash2.py:67-72 -- the float focal / principal point are TRUNCATED by the
  assignment into an integer tensor)  ->  `intrinsics_matrix`
* `bounds_model.npy`: float32 (2,3) = [min_bound, max_bound] (train_hash2.py:115, read at nerf2mesh.py:27-28)
* `density_grid_w_rgb.npy`: float32 (res,res,res,4) = [rgb, density] over np.meshgrid(x,y,z) in its default 'xy'
  indexing, i.e. axis 0 = y, axis 1 = x, axis 2 = z (nerf2mesh.py:30-37,86-87)
* checkpoints: `{name}_Nerf_hash.pth` / `{name}_encoder_hash.pth` = torch.save(state_dict) (train_hash2.py:299-300, loaded
  at :129-133 and nerf2mesh.py:60-63); MLP keys carry the `module.` prefix of nn.DataParallel
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------------------------
# scene description
# ---------------------------------------------------------------------------------------------------------------
def _frame_file(json_path: str, file_path: str, add_png: bool) -> str:
    """dataset.py:22,35 / dataset_new.py:22,36: directory of the json + everything after the FIRST '.' of file_path
    ('./train/r_0' -> '/train/r_0'), plus '.png' for the Blender flavour."""
    return json_path[:json_path.rfind("/")] + file_path[file_path.find(".") + 1:] + (".png" if add_png else "")


def read_transforms(json_path: str) -> Dict:
    """Parses either flavour.  Returns H, W, focal1, focal2, cx, cy exactly as the reference's dataset objects expose
    them (Blender: focal = W / (2 tan(camera_angle_x / 2)) as a float32 tensor, cx = W/2, cy = H/2; new: the json's
    numbers), `files` (absolute image paths), `c2w` (V,4,4) float32 and `extra` (per-frame rotation / sharpness)."""
    if not os.path.exists(json_path):
        raise FileNotFoundError(f"The path {json_path} does not exist")
    with open(json_path, "r") as f:
        data = json.load(f)
    frames = data["frames"]
    new = "fl_x" in data
    files = [_frame_file(json_path, fr["file_path"], add_png=not new) for fr in frames]
    c2w = torch.stack([torch.Tensor(fr["transform_matrix"]) for fr in frames]) if frames else torch.zeros(0, 4, 4)
    out = {"flavour": "new" if new else "blender", "files": files, "c2w": c2w,
           "camera_angle_x": torch.tensor(data["camera_angle_x"]) if "camera_angle_x" in data else None}
    if new:
        out.update(H=data["h"], W=data["w"], focal1=data["fl_x"], focal2=data["fl_y"], cx=data["cx"], cy=data["cy"],
                   extra=[fr.get("sharpness") for fr in frames])
    else:
        import cv2
        img = cv2.imread(files[0])
        if img is None:
            raise FileNotFoundError(f"The file {files[0]} does not exist")
        H, W, _ = img.shape
        focal = W / (2 * torch.tan(out["camera_angle_x"] / 2))          # dataset.py:26 (float32 tensor arithmetic)
        out.update(H=H, W=W, focal1=focal, focal2=focal, cx=W / 2, cy=H / 2, extra=[fr.get("rotation") for fr in frames])
    return out


def intrinsics_matrix(meta: Dict) -> torch.Tensor:
    """train_hash2.py:67-72: an int64 identity whose entries are overwritten by the (float) focal lengths and principal
    point -- the assignment truncates toward zero, and get_od divides by these integers."""
    K = torch.from_numpy(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]]))
    K[0, 0] = meta["focal1"]
    K[1, 1] = meta["focal2"]
    K[0, 2] = meta["cx"]
    K[1, 2] = meta["cy"]
    return K


def read_image_u8(path: str) -> torch.Tensor:
    """(H,W,3) uint8 RGB -- cv2.imread + BGR->RGB as dataset.py:37-38.  torchvision's ToTensor of that array is exactly
    this / 255 as float32, permuted to (3,H,W); rays.DeviceRayDataset keeps the uint8 and divides on the device."""
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(f"The file {path} does not exist")
    return torch.from_numpy(cv2.cvtColor(img, cv2.COLOR_BGR2RGB))


def load_scene(json_path: str, limit: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Dict]:
    """(images (V,H,W,3) uint8, c2w (V,4,4) float32, K int64 3x3, meta): everything rays.DeviceRayDataset needs."""
    meta = read_transforms(json_path)
    files = meta["files"][:limit] if limit is not None else meta["files"]
    images = torch.stack([read_image_u8(f) for f in files])
    return images, meta["c2w"][: len(files)], intrinsics_matrix(meta), meta


# ---------------------------------------------------------------------------------------------------------------
# bounds / density grid / checkpoints
# ---------------------------------------------------------------------------------------------------------------
def save_bounds(path: str, min_bound, max_bound) -> None:
    """train_hash2.py:115: np.save('bounds_model.npy', torch.stack([min_bound, max_bound]).numpy())."""
    np.save(path, torch.stack([torch.as_tensor(min_bound).detach().cpu(), torch.as_tensor(max_bound).detach().cpu()]).numpy())


def load_bounds(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """nerf2mesh.py:27-28: (min_bound, max_bound)."""
    b = np.load(path)
    if b.shape != (2, 3):
        raise ValueError(f"{path}: expected a (2,3) array [min_bound, max_bound], got {b.shape}")
    return b[0], b[1]


def save_density_grid(path: str, grid) -> None:
    """nerf2mesh.py:86-87: (res,res,res,4) float32 [rgb, density]."""
    g = grid.detach().cpu().numpy() if torch.is_tensor(grid) else np.asarray(grid)
    if g.ndim != 4 or g.shape[-1] != 4 or not (g.shape[0] == g.shape[1] == g.shape[2]):
        raise ValueError(f"expected (res,res,res,4), got {g.shape}")
    np.save(path, g.astype(np.float32, copy=False))


def load_density_grid(path: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """nerf2mesh.py:88-95: returns (out_grid (res,res,res,4), density_grid = out_grid[..., -1].clone())."""
    out = torch.tensor(np.load(path))
    if out.dim() != 4 or out.shape[-1] != 4:
        raise ValueError(f"{path}: expected (res,res,res,4), got {tuple(out.shape)}")
    return out, out[..., -1].clone()


def grid_axes(min_bound, max_bound, res: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The coordinates behind density[i,j,k] (nerf2mesh.py:30-37): i indexes y, j indexes x, k indexes z."""
    x = np.linspace(min_bound[0], max_bound[0], res)
    y = np.linspace(min_bound[1], max_bound[1], res)
    z = np.linspace(min_bound[2], max_bound[2], res)
    return y, x, z


def checkpoint_paths(name: str) -> Tuple[str, str]:
    return name + "_Nerf_hash.pth", name + "_encoder_hash.pth"


def save_checkpoint(name: str, nerf, encoder) -> None:
    """train_hash2.py:299-300."""
    n, e = checkpoint_paths(name)
    torch.save(nerf.state_dict(), n)
    torch.save(encoder.state_dict(), e)


def load_checkpoint(name: str, nerf, encoder, map_location=None) -> None:
    """train_hash2.py:129-133 / nerf2mesh.py:60-63.  Accepts MLP checkpoints saved with or without the `module.` prefix of
    nn.DataParallel, whichever way `nerf` itself is wrapped."""
    n, e = checkpoint_paths(name)
    sd = torch.load(n, map_location=map_location)
    want_prefix = isinstance(nerf, torch.nn.DataParallel)
    has_prefix = all(k.startswith("module.") for k in sd)
    if want_prefix and not has_prefix:
        sd = {"module." + k: v for k, v in sd.items()}
    elif has_prefix and not want_prefix:
        sd = {k[len("module."):]: v for k, v in sd.items()}
    nerf.load_state_dict(sd)
    encoder.load_state_dict(torch.load(e, map_location=map_location))
