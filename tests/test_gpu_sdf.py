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
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
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


ROUTES = ("native", "generic", "composed")


@pytest.fixture(params=ROUTES)
def route(request):
    """native: Volume_Renderer's own SDF route (field as in NeRF mode -> hbr_composite_sdf_* on the packed output -> the
    eikonal stencil); generic: the reference's data flow with calc_color on the SDF kernels; composed: the tensor
    expressions of the reference on the device (the A/B the kernels replace)."""
    import human_body_reconstruction_b200 as h
    old = h.helper.SDF_KERNELS
    h.helper.SDF_KERNELS = request.param != "composed"
    yield request.param
    h.helper.SDF_KERNELS = old


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
    # the stencil form (one encoder + density-head pass over the six clamped positions) gives the same differences
    with torch.no_grad():
        norm, grads6 = mlp.eikonal_norms(pts, encoder=enc, with_grads=True)
    assert float((norm.cpu() - g["norm"]).abs().max()) < 2e-3
    assert float((grads6 - grads).abs().max()) < 1e-3
    assert torch.allclose(norm, h.helper.eikonal_value(grads6), rtol=1e-6, atol=0)


def test_sdf_vol_render_loss_and_gradients_match_reference(route):
    g = load_golden("sdf.npz")
    h, enc, mlp, var, vr = _build(g)
    vr.sdf_native = route == "native"
    calls0 = dict(h._lib.STATS.calls)
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
    used = {k: v - calls0.get(k, 0) for k, v in h._lib.STATS.calls.items() if "sdf" in k and v > calls0.get(k, 0)}
    if route == "composed":
        assert not used, used
    else:                                                                # one launch each, forward and backward
        assert used == {"hbr_composite_sdf_fwd": 1, "hbr_composite_sdf_bwd": 1, "hbr_sdf_stencil_points": 1,
                        "hbr_sdf_eikonal_fwd": 1, "hbr_sdf_eikonal_bwd": 1}, used


def test_sdf_hierarchical_fails_like_the_reference():
    """The reference's fine pass calls calc_color without the sample positions (vol_renderer.py:242), so SDF mode with
    hierarchical=True dies in finite_difference_normals_approximator(None) on `x.device` (test_hash.py:90): an
    AttributeError; same here (no silent fallback)."""
    g = load_golden("sdf.npz")
    h, enc, mlp, var, vr = _build(g)
    with pytest.raises(AttributeError):
        vr.vol_render(mlp, g["rays_d"].to(DEV), g["rays_o"].to(DEV), num_samples=8, update_mask=False,
                      dir_norm=g["dir_norm"].to(DEV), hierarchical=True)


# ---- kernel level: hbr_composite_sdf_fwd/bwd against the oracle's restatement in float64 ---------------------------------
def _sdf_case(kind, R, S, seed):
    gen = torch.Generator().manual_seed(seed)
    if kind == "uniform":                   # independent values in (-1, 1): large alphas
        sdf, b = torch.rand(R, S, generator=gen) * 2 - 1, 0.5
    elif kind == "walk":                    # a smooth profile (what a trained field gives): small alphas; steps kept away from
        step = torch.randn(R, S, generator=gen)          # 0 so that relu(1 - phi'/phi) is not evaluated AT its kink
        sdf, b = torch.cumsum(torch.sign(step) * (1e-3 + 0.05 * step.abs()), -1), 0.5
    else:                                   # wide range with the -10 clamp of helper.py:76 hit
        sdf, b = (torch.rand(R, S, generator=gen) * 2 - 1) * 8, 1.3
        sdf[0, 0] = -12.0
        sdf[-1, -1] = -10.5
    rgb = torch.rand(R, S, 3, generator=gen)
    gC = torch.randn(R, 3, generator=gen)
    gw = torch.randn(R, S, generator=gen) * 0.1
    return rgb, sdf, b, gC, gw


@pytest.mark.parametrize("S", [1, 2, 31, 33, 128, 200, 700])
@pytest.mark.parametrize("kind", ["uniform", "walk", "clamped"])
def test_composite_sdf_kernels_match_oracle(kind, S):
    """Tolerances (norm-wise, against float64 on the same fp32 inputs): colours / weights / d rgb 5e-5 -- alpha = 1 - phi'/phi
    cancels: the same closed form evaluated in fp32 by torch on the CPU is 5e-6 away on the smooth profile (3e-7 on the
    others) --, d sdf 1e-4 (fp32 torch: 1e-6), dL/db 2e-4."""
    from human_body_reconstruction_b200 import ops
    from oracle import port
    R = 37
    rgb, sdf, b, gC, gw = _sdf_case(kind, R, S, 100 + S)
    r64, s64 = rgb.double().requires_grad_(), sdf.double().requires_grad_()
    b64 = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    C_ref, w_ref = port.composite_sdf(r64, s64, b64)
    ((C_ref * gC.double()).sum() + (w_ref[..., 0] * gw.double()).sum()).backward()

    rg, sg = rgb.to(DEV).requires_grad_(), sdf.to(DEV).requires_grad_()
    bg = torch.tensor(b, device=DEV, requires_grad=True)
    sdf_before = sg.detach().clone()
    C, w = ops.CompositeSdf.apply(rg, sg, bg, False)
    assert torch.equal(sg.detach(), sdf_before)                          # the clamp is applied inside, not to the caller's tensor
    assert _rel(C, C_ref.detach()) < 5e-5 and _rel(w, w_ref.detach()[..., 0]) < 5e-5
    ((C * gC.to(DEV)).sum() + (w * gw.to(DEV)).sum()).backward()
    assert _rel(rg.grad, r64.grad) < 5e-5
    assert _rel(sg.grad, s64.grad) < 1e-4
    assert abs(float(bg.grad) - float(b64.grad)) < 2e-4 * abs(float(b64.grad)) + 2e-5
    if kind == "clamped":
        assert float(sg.grad[0, 0]) == 0.0 and float(sg.grad[-1, -1]) == 0.0
    # packed layout (the MLP's (R*S,4) output, strides 4): same kernel, column views
    out4 = torch.cat((rgb.reshape(-1, 3), sdf.reshape(-1, 1)), dim=-1).detach().to(DEV).requires_grad_()
    C2, w2 = ops.CompositeSdfPacked.apply(out4, bg.detach(), R, S, False)
    assert torch.equal(C2, C) and torch.equal(w2, w)
    (C2 * gC.to(DEV)).sum().backward()                                   # weights unused: no gradient tensor for them
    assert out4.grad.shape == (R * S, 4) and torch.isfinite(out4.grad).all()


def test_composite_sdf_from_density_matches_the_composed_form():
    """from_density: column 3 holds the density head's LeakyReLU output; the kernel forms 2*sigmoid(pre-activation) - 1
    itself (test_hash.py:59-60) and hands back the gradient with respect to the LeakyReLU output."""
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200 import ops
    from oracle import port
    R, S = 29, 96
    gen = torch.Generator().manual_seed(5)
    dens = torch.randn(R * S, generator=gen) * 1.5
    dens = torch.where(dens > 0, dens, dens * 0.01)                      # what LeakyReLU(0.01) leaves
    rgb = torch.rand(R * S, 3, generator=gen)
    gC = torch.randn(R, 3, generator=gen)
    d64 = dens.double().requires_grad_()
    raw = torch.where(d64 > 0, d64, d64 * 100.0)
    b64 = torch.tensor(0.8, dtype=torch.float64, requires_grad=True)
    C_ref, _ = port.composite_sdf(rgb.double().reshape(R, S, 3), (2 * torch.sigmoid(raw) - 1).reshape(R, S), b64)
    (C_ref * gC.double()).sum().backward()
    out4 = torch.cat((rgb, dens[:, None]), dim=-1).detach().to(DEV).requires_grad_()
    bg = torch.tensor(0.8, device=DEV, requires_grad=True)
    C, _ = ops.CompositeSdfPacked.apply(out4, bg, R, S, True)
    assert _rel(C, C_ref.detach()) < 2e-5
    (C * gC.to(DEV)).sum().backward()
    assert _rel(out4.grad[:, 3], d64.grad) < 1e-4
    assert abs(float(bg.grad) - float(b64.grad)) < 2e-4 * abs(float(b64.grad)) + 2e-5


def test_eikonal_stencil_kernels():
    """hbr_sdf_stencil_points: bit-identical to (x +- eps e).clamp(lo, hi) (test_hash.py:91-102); hbr_sdf_eikonal_fwd/bwd
    against float64 autograd of 0.5 (s+ - s-) / eps -> norm on the same inputs (1e-5: values well apart; the realistic
    nearly-equal pairs are covered by the fixture tests above at 2e-3 absolute)."""
    from human_body_reconstruction_b200 import ops
    gen = torch.Generator().manual_seed(9)
    n, eps = 1000, 0.0005
    lo, hi = [-1.0, -0.5, -2.0], [1.0, 0.75, 2.0]
    x = (torch.rand(n, 3, generator=gen) * 2 - 1) * torch.tensor([1.2, 0.9, 2.2])
    x[0] = torch.tensor([1.0, 0.75, -2.0])                               # on the faces: one side of the stencil is clamped
    xd = x.to(DEV)
    pts = ops.sdf_stencil_points(xd, eps, lo, hi)
    lo_t, hi_t = torch.tensor(lo, device=DEV), torch.tensor(hi, device=DEV)
    for axis in range(3):
        e = torch.zeros(1, 3, device=DEV)
        e[0, axis] = eps
        assert torch.equal(pts[2 * axis], (xd + e).clamp(lo_t, hi_t))
        assert torch.equal(pts[2 * axis + 1], (xd - e).clamp(lo_t, hi_t))
    dens = torch.randn(6, n, generator=gen)
    dens = torch.where(dens > 0, dens, dens * 0.01)
    d64 = dens.double().requires_grad_()
    s64 = 2 * torch.sigmoid(torch.where(d64 > 0, d64, d64 * 100.0)) - 1
    g64 = torch.stack([0.5 * (s64[2 * a] - s64[2 * a + 1]) / eps for a in range(3)], dim=-1)
    n64 = torch.sqrt((g64 ** 2).sum(-1))
    gn = torch.randn(n, generator=gen)
    (n64 * gn.double()).sum().backward()
    dg = dens.to(DEV).requires_grad_()
    norm, grads = ops.SdfEikonal.apply(dg, eps)
    assert _rel(norm, n64.detach()) < 1e-5 and _rel(grads, g64.detach()) < 1e-5
    (norm * gn.to(DEV)).sum().backward()
    assert _rel(dg.grad, d64.grad) < 1e-5


def test_native_sdf_route_under_autocast_runs_the_tensor_core_field():
    """Under autocast the native route takes the NeRF-mode field kernels (tcgen05 MLP) and only then the SDF compositor;
    the eikonal norms always come from the fp32 kernels (bit-identical).  The SDF alphas are DIFFERENCES of neighbouring
    field values (1 - phi'/phi), so 16-bit operand rounding (3e-4 on the field outputs) is amplified in the colours:
    bound 5e-2 here."""
    g = load_golden("sdf.npz")
    h, enc, mlp, var, vr = _build(g)
    S = g["t"].shape[0]
    args = (mlp, g["rays_d"].to(DEV), g["rays_o"].to(DEV))
    kw = dict(num_samples=S, t=g["t"].to(DEV), update_mask=False, dir_norm=g["dir_norm"].to(DEV), hierarchical=False)
    with torch.no_grad():
        Cr, _, norm = vr.vol_render(*args, **kw)
        with torch.autocast("cuda", dtype=torch.float16):
            Cr16, _, norm16 = vr.vol_render(*args, **kw)
    assert _rel(Cr16, Cr) < 5e-2, _rel(Cr16, Cr)
    assert torch.equal(norm16, norm)


def test_density_only_backward_equals_the_density_column_of_the_full_field():
    """hbr_mlp_bwd_f32 with dirs == NULL (the sigma-net pass of the eikonal stencil): d(features) bit-identical to the full
    field's backward fed a gradient on the density column only (the per-point chain is the same arithmetic), density-head
    parameter gradients equal up to the order of the atomic accumulation, colour-net gradients exactly zero."""
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200.test_hash import _DensityFn
    torch.manual_seed(3)
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24).to(DEV)
    n = 5000
    feat = torch.randn(n, 32, device=DEV)
    gd = torch.randn(n, 1, device=DEV)
    f1 = feat.clone().requires_grad_()
    d1 = _DensityFn.apply(f1, mlp, *mlp._ordered())
    d1.backward(gd)
    g1 = {k: v.grad.clone() for k, v in mlp.named_parameters()}
    for p in mlp.parameters():
        p.grad = None
    f2 = feat.clone().requires_grad_()
    d2 = mlp.field(f2, torch.zeros(n, 24, device=DEV), 1, use_tc=False, raw=True)[:, 3:4]
    d2.backward(gd)
    assert torch.equal(d1, d2) and torch.equal(f1.grad, f2.grad)
    for k, v in mlp.named_parameters():
        if k.startswith("sig_model"):
            assert _rel(g1[k], v.grad) < 1e-5, k
        else:
            assert float(g1[k].abs().max()) == 0.0 and float(v.grad.abs().max()) == 0.0, k
