"""human_body_reconstruction_b200: B200-native (sm_100a) implementation of the NeRF hot path of
RishabhSri14/Human-Body-Reconstruction -- multiresolution hash encoder, field MLP, ray sampling +
alpha compositing, density grid + marching cubes -- behind the reference's own module/class names.

    from human_body_reconstruction_b200 import HashEncoder, PositionalEncoder, MLP_3D, Volume_Renderer

The `dropin/` directory at the repo root re-exports these under the reference's flat module names
(hash_encoding, encoder, test_hash, helper, vol_renderer, tmp_encoder) so train_hash2.py / nerf2mesh.py
run unchanged with PYTHONPATH=dropin.
"""
from .encoder import PositionalEncoder
from .hash_encoding import HashEncoder
from .test_hash import MLP_3D
from .vol_renderer import Volume_Renderer
from . import dist, formats, graph, helper, mesh, ops, optim, peer, rays  # noqa: F401
from .rays import DeviceRayDataset

__all__ = ["HashEncoder", "PositionalEncoder", "MLP_3D", "Volume_Renderer", "DeviceRayDataset", "dist", "formats", "graph", "helper", "mesh",
           "ops", "optim", "peer", "rays"]
