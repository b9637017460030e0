"""Stand-in: the reference imports matplotlib.pyplot / matplotlib.lines.Line2D (helper.py:7-8, train_hash2.py:6,
nerf2mesh.py:3); only plot_grad_flow (--plot_grads) would use them."""
from . import lines, pyplot  # noqa: F401
