"""Reference-named module: `helper` of RishabhSri14/Human-Body-Reconstruction, served by the B200 package.
Put this directory first on PYTHONPATH and the reference's train_hash2.py / nerf2mesh.py import it unchanged."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_body_reconstruction_b200.helper import *  # noqa: F401,F403
import torch, numpy as np, time  # noqa: F401,E401
from torch import nn  # noqa: F401
from typing import Tuple, Optional  # noqa: F401
