"""Reference-named module: `test_hash` of RishabhSri14/Human-Body-Reconstruction, served by the B200 package.
Put this directory first on PYTHONPATH and the reference's train_hash2.py / nerf2mesh.py import it unchanged."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_body_reconstruction_b200.test_hash import *  # noqa: F401,F403
from human_body_reconstruction_b200.hash_encoding import HashEncoder  # noqa: F401  (test_hash.py:2)
import torch, numpy as np  # noqa: F401,E401
