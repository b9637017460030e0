"""SDF mode on the CUDA path against the fixture produced by the reference's
Volume_Renderer(use_sdf=True) (tests/golden/sdf.npz, oracle/make_golden.py: gold_sdf).
Tolerances: colours 1e-4 relative to the largest colour; the eikonal norms are central differences with eps = 5e-4, which
amplify fp32 rounding of the SDF value by 1/(2 eps) = 1000, hence 2e-3 absolute there and 2e-3 norm-wise on the gradients
that flow through them."""
import pytest
import torch

from conftest import load_golden, mlp_params

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _build(g):
    import human_body_reconstruction_b200 as h
    L, T, F = g["tables"].shape
    enc = h.HashEncoder(N_min=16, N_max=512.0, L=L, F=F, T=T, dim=3, mu=g["mu"].to(DEV), sigma=g["sigma"].to(DEV))
    enc.load_state_dict({f"Embedding_list.{i}.weight": g["tables"][i] for i in range(L)})
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=L, F=F, d_view=24, use_sdf=True, max_bound=g["max_bound"], min_bound=g["min_bound"])
    mlp.load_state_dict(mlp_params(g, "mlp__"))
    var = h.helper.VarModel().to(DEV)
    with torch.no_grad():
        var.b.fill_(float(g["b"]))
    enc, mlp = enc.to(DEV), mlp.to(DEV)
    vr = h.Volume_Renderer(H=8, W=8, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device=DEV, Pos_encode=enc,
                           Dir_encode=h.PositionalEncoder(3, 4), max_dim=64, sigma_val=g["sigma"], mu=g["mu"], use_sdf=True,
                           var_model=var)
    return h, enc, mlp, var, vr


def test_forward_sdf_and_normals_match_reference():
    g = load_golden("sdf.npz")
    h, enc, mlp, var, vr = _build(g)
    pts = (g["rays_o"][:, None, :] + g["rays_d"][:, None, :] * g["t"][None, :, None]).reshape(-1, 3).to(DEV)
    with torch.no_grad():
        sdf = mlp.forward_sdf(pts, encoder=enc)
        grads = mlp.finite_difference_normals_approximator(pts, encoder=enc)
    assert torch.allclose(sdf.cpu(), g["sdf"], rtol=1e-5, atol=1e-7)
    assert float((h.helper.eikonal_value(grads).cpu() - g["norm"]).abs().max()) < 2e-3
    # the model's own forward in SDF mode: density column in (-1, 1), equal to forward_sdf of the same features
    feat = enc(pts)
    out = mlp(feat, torch.zeros(pts.shape[0], 24, device=DEV))
    assert torch.allclose(out[:, 3:4], sdf, rtol=1e-5, atol=1e-7) and float(out[:, 3].abs().max()) < 1


def test_sdf_vol_render_loss_and_gradients_match_reference():
    g = load_golden("sdf.npz")
    h, enc, mlp, var, vr = _build(g)
    model = torch.nn.DataParallel(mlp, device_ids=[0])                   # train_hash2.py:127; helper.py:87 uses .module
    S = g["t"].shape[0]
    Cr, Cf, norm = vr.vol_render(model, g["rays_d"].to(DEV), g["rays_o"].to(DEV), num_samples=S, t=g["t"].to(DEV),
                                 update_mask=False, dir_norm=g["dir_norm"].to(DEV), hierarchical=False)
    assert Cf is Cr and norm.shape == g["norm"].shape
    assert float((Cr.detach().cpu() - g["Cr"]).abs().max()) < 1e-4 * float(g["Cr"].abs().max())
    assert float((norm.detach().cpu() - g["norm"]).abs().max()) < 2e-3
    gt = g["gt"].to(DEV)
    loss = (torch.nn.functional.mse_loss(Cr, gt) + torch.nn.functional.mse_loss(Cf, gt)
            + 0.1 * h.helper.eikonal_loss(norm))                         # train_hash2.py:221-224
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * float(g["loss"])
    loss.backward()
    assert abs(float(var.b.grad) - float(g["grad_b"])) < 2e-3 * abs(float(g["grad_b"])) + 1e-9
    assert _rel(torch.stack([e.weight.grad for e in enc.Embedding_list]), g["dtables"]) < 2e-3
    for k, v in mlp.named_parameters():
        assert _rel(v.grad, g["grad__" + k.replace(".", "__")]) < 2e-3, k


def test_sdf_hierarchical_fails_like_the_reference():
    """The reference's fine pass calls calc_color without the sample positions (vol_renderer.py:242), so SDF mode with
    hierarchical=True dies in finite_difference_normals_approximator(None) on `x.device` (test_hash.py:90): an
    AttributeError; same here (no silent fallback)."""
    g = load_golden("sdf.npz")
    h, enc, mlp, var, vr = _build(g)
    with pytest.raises(AttributeError):
        vr.vol_render(mlp, g["rays_d"].to(DEV), g["rays_o"].to(DEV), num_samples=8, update_mask=False,
                      dir_norm=g["dir_norm"].to(DEV), hierarchical=True)
