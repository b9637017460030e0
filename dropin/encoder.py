"""Reference-named module: `encoder` of RishabhSri14/Human-Body-Reconstruction, served by the B200 package.
Put this directory first on PYTHONPATH and the reference's train_hash2.py / nerf2mesh.py import it unchanged."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_body_reconstruction_b200.encoder import *  # noqa: F401,F403
import torch, torch.nn as nn, numpy as np  # noqa: F401,E401
import torch.nn.functional as F  # noqa: F401
