"""-m gpu: the density-grid query of nerf2mesh.py:69-87 on the split-TF32 tensor-core density head
(hbr_mlp_density_tf32x3) against the fp32 CUDA-core kernel, a float64 evaluation and the reference's own fixture.
The path's bar is fp32 parity (1e-5): the reference evaluates the grid without autocast."""
import pytest
import torch

from conftest import load_golden, mlp_params
from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def density_f64(feat, p):
    """sig_model in float64 (test_hash.py:52-62): the yardstick both kernels are measured against."""
    x = feat.double().cpu()
    for i in (0, 2):
        x = torch.relu(x @ p[f"sig_model.{i}.weight"].double().T + p[f"sig_model.{i}.bias"].double())
    o = x @ p["sig_model.4.weight"].double().T + p["sig_model.4.bias"].double()
    d = o[:, 0]
    return torch.where(d > 0, d, 0.01 * d)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 300_001])
@pytest.mark.parametrize("scale", [1.0, 30.0])
def test_density_tf32x3_matches_fp32(n, scale):
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200 import ops
    torch.manual_seed(n)
    p = port.mlp_init(seed=3)
    m = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=torch.ones(3), min_bound=-torch.ones(3))
    m.load_state_dict(p)
    m = m.to(DEV)
    feat = (scale * torch.randn(n, 32)).to(DEV)
    simt, _ = ops.mlp_fwd_f32(feat, None, 1, m._flat_params(), m._dims(), False)
    tc = ops.mlp_density_tf32x3(feat, m._flat_params(), m._dims())
    ref = density_f64(feat, p)
    e_simt, e_tc = rel(simt[:, 0], ref), rel(tc, ref)
    print(f"n={n} scale={scale}: fp32 CUDA cores {e_simt:.2e}, split-TF32 tensor cores {e_tc:.2e} (vs float64)")
    assert e_tc < 1e-5 and rel(tc, simt[:, 0]) < 1e-5
    # element-wise, relative to the magnitude of the terms of each dot product (outputs cancel)
    assert ((tc.double().cpu() - ref).abs() <= 1e-5 * ref.abs() + 1e-5 * ref.abs().max()).all()


def test_density_grid_default_is_tensor_core_and_matches_reference_fixture():
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200 import _lib
    from test_gpu_parity import make_encoder, make_mlp
    g = load_golden("grid.npz")
    enc, mlp = make_encoder(g), make_mlp(mlp_params(g, "mlp__"))
    res = int(g["res"])
    mn, mx = g["min_bound"].double().tolist(), g["max_bound"].double().tolist()
    _lib.STATS.reset()
    dens = h.mesh.density_grid(enc, mlp, None, mn, mx, res)
    exact = h.mesh.density_grid(enc, mlp, None, mn, mx, res, cuda_core_mlp=True)
    want = g["out"][..., 3]
    print(f"grid fixture: tensor cores {rel(dens, want):.2e}, CUDA cores {rel(exact, want):.2e}")
    assert rel(dens, want) < 1e-5 and rel(exact, want) < 1e-5
    assert torch.allclose(dens.cpu(), want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))
    # marching cubes sees the same surface: identical vertex / triangle counts at an iso level inside the value range
    iso = float(want.median())
    assert h.mesh.marching_cubes_counts(dens, iso) == h.mesh.marching_cubes_counts(exact, iso)
