"""CPU: the product never falls back.  Without a CUDA device every entry of the package raises (RuntimeError naming the
missing CUDA path) instead of computing on the host, and nothing under human_body_reconstruction_b200/ or dropin/ imports,
links or executes oracle/ (test infrastructure only)."""
import os
import re

import pytest
import torch

from conftest import ROOT

needs_no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a host without CUDA")


@needs_no_gpu
def test_modules_raise_without_cuda():
    import human_body_reconstruction_b200 as h
    enc = h.HashEncoder(N_min=16, N_max=64.0, L=2, F=2, T=64, dim=3, mu=torch.zeros(3), sigma=torch.tensor(1.0), device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.rand(4, 3))
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24)
    with pytest.raises(RuntimeError, match="CUDA"):
        mlp(torch.rand(4, 32), torch.rand(4, 24))
    with pytest.raises(RuntimeError, match="CUDA"):
        h.helper.calc_color(torch.linspace(2, 6, 4), torch.rand(2, 4, 3), torch.rand(2, 4), torch.ones(2, 1))
    with pytest.raises(RuntimeError, match="CUDA"):                      # SDF mode (8f row 4): compositor, stencil, eikonal
        h.helper.calc_color(torch.linspace(2, 6, 4), torch.rand(2, 4, 3), torch.rand(2, 4), torch.ones(2, 1), use_sdf=True,
                            var_model=h.helper.VarModel(), rays=torch.rand(8, 3), model=mlp, encoder=enc)
    with pytest.raises(RuntimeError, match="CUDA"):
        mlp.eikonal_norms(torch.rand(4, 3), encoder=enc)
    with pytest.raises(RuntimeError, match="CUDA"):
        h.ops.CompositeSdf.apply(torch.rand(2, 4, 3), torch.rand(2, 4), torch.tensor(0.5), False)
    with pytest.raises(RuntimeError, match="CUDA"):
        h.ops.sdf_stencil_points(torch.rand(4, 3), 5e-4, [-1.0] * 3, [1.0] * 3)
    vr = h.Volume_Renderer(H=4, W=4, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device="cpu", Pos_encode=enc,
                           Dir_encode=h.PositionalEncoder(3, 4), max_dim=16, sigma_val=torch.tensor(1.0), mu=torch.zeros(3))
    with pytest.raises(RuntimeError, match="CUDA"):
        vr.vol_render(mlp, torch.rand(2, 3), torch.rand(2, 3), num_samples=4, hierarchical=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        h.ops.hash_encode_fwd(torch.rand(4, 3), torch.rand(2, 64, 2), enc._geom())
    with pytest.raises(RuntimeError, match="CUDA"):
        h.DeviceRayDataset(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), torch.eye(4)[None], torch.eye(3), device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        h.peer.PeerRegion(16)
    prm = torch.nn.Parameter(torch.zeros(4))
    prm.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        h.optim.FusedAdam([prm]).step()
    assert torch.equal(prm.detach(), torch.zeros(4))                     # nothing was updated on the host


def test_product_never_touches_the_oracle():
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.\.?oracle)|oracle[./]port|ref_loader", re.M)
    offenders = []
    for top in ("human_body_reconstruction_b200", "dropin"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(d, f), errors="ignore").read()
                    if pat.search(src):
                        offenders.append(os.path.join(d, f))
    assert not offenders, offenders
