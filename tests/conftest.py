import contextlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def mlp_params(g, prefix=""):
    """state-dict style dict from a fixture whose keys use '__' for '.'."""
    out = {}
    for k, v in g.items():
        if k.startswith(prefix) and ("sig_model" in k or "col_model" in k) and "grad" not in k:
            out[k[len(prefix):].replace("__", ".")] = v
    return out


@pytest.fixture
def golden():
    return load_golden


@contextlib.contextmanager
def capture_fine_sampling():
    """Records what Volume_Renderer hands to / gets from hierarchical_sampling: rec["w"] (coarse weights before the in-place
    clamp, (R,S)) and rec["t_fine"] ((R,2S))."""
    from human_body_reconstruction_b200 import vol_renderer as vrm
    rec = {}
    orig = vrm.hierarchical_sampling

    def wrapped(*a, **k):
        w = k["weights"]
        rec["w"] = (w.squeeze(-1) if w.dim() == 3 else w).detach().clone()
        rays, tf = orig(*a, **k)
        rec["t_fine"] = tf.detach().clone()
        return rays, tf

    vrm.hierarchical_sampling = wrapped
    try:
        yield rec
    finally:
        vrm.hierarchical_sampling = orig
