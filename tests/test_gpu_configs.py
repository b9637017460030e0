"""GPU tests (-m gpu) at the sizes of BASELINE.json's other configurations, through size-independent properties
(the CPU oracle is only run where it finishes in seconds):
  configs[3]  nerf2mesh density-grid query at 512^3 + marching cubes, sharded by slab;
  configs[4]  human-reconstruction configuration: T = 2^22, 256 samples/ray, hierarchical (768 points/ray)."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda"
MU, MAXB = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
SIGMA = ((MAXB - MU) ** 2).sum().sqrt()


def build(T, scale=5e3, seed=0):
    import human_body_reconstruction_b200 as h
    torch.manual_seed(seed)
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=MU.to(DEV), sigma=SIGMA.to(DEV))
    with torch.no_grad():
        for e in enc.Embedding_list:
            e.weight.mul_(scale)
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=MAXB, min_bound=MU)
    enc, mlp = enc.to(DEV), mlp.to(DEV)
    vr = h.Volume_Renderer(H=8, W=8, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device=DEV,
                           Pos_encode=enc, Dir_encode=h.PositionalEncoder(3, 4), max_dim=64, sigma_val=SIGMA, mu=MU)
    return h, enc, mlp, vr


def rays(R, seed=1):
    g = torch.Generator().manual_seed(seed)
    ro = torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.05 * torch.randn(R, 3, generator=g)
    rd = torch.nn.functional.normalize(-ro + 0.5 * torch.randn(R, 3, generator=g), dim=-1)
    return ro, rd, 1 + 0.2 * torch.rand(R, 1, generator=g), torch.rand(R, 3, generator=g)


def test_density_grid_512_slabs_and_marching_cubes():
    """configs[3]: 512^3 = 134 217 728 grid points.  Slabs evaluated separately (the multi-GPU decomposition) are
    bit-identical to the same planes of a whole-grid pass; slab-owned marching-cubes counts add up to the whole grid's
    and the welded vertex count equals the number of iso-crossing grid edges (torch reduction on the device)."""
    h, enc, mlp, _ = build(2 ** 19, scale=2e5)                 # large table values: densities straddle the iso level
    res = 512
    mn, mx = MU.double().tolist(), MAXB.double().tolist()
    dens = h.mesh.density_grid(enc, mlp, None, mn, mx, res)
    assert dens.shape == (res, res, res) and bool(torch.isfinite(dens).all())
    for r in range(8):                                          # z-slab sharding over 8 ranks, checked on 3 of them
        if r in (0, 3, 7):
            i0, i1 = h.dist.slab_range(res, r, 8) if hasattr(h, "dist") else (r * 64, (r + 1) * 64)
            slab = h.mesh.density_grid(enc, mlp, None, mn, mx, res, i_begin=i0, i_end=i1)
            assert torch.equal(slab, dens[i0:i1])
    iso = float(dens.median())
    inside = dens < iso
    want = int((inside[1:] != inside[:-1]).sum() + (inside[:, 1:] != inside[:, :-1]).sum() + (inside[:, :, 1:] != inside[:, :, :-1]).sum())
    nv, nt = h.mesh.marching_cubes_counts(dens, iso)
    assert nv == want and nt > 0
    parts = [h.mesh.marching_cubes_counts(dens, iso, r * 64, (r + 1) * 64) for r in range(8)]
    assert sum(p[0] for p in parts) == nv and sum(p[1] for p in parts) == nt
    # a corner of the grid against the CPU oracle (the whole grid would take the oracle minutes)
    sub = dens[:12, :12, :12].cpu().numpy()
    assert h.mesh.marching_cubes_counts(dens[:12, :12, :12].contiguous(), iso)[0] == port.mc_crossing_edges(sub, iso)


def test_human_config_small_vs_oracle():
    """configs[4] shape at a ray count the oracle handles: T = 2^22, 256 coarse + 512 fine samples/ray, fp32 path."""
    h, enc, mlp, vr = build(2 ** 22)
    R, S = 6, 256
    ro, rd, dn, gt = rays(R)
    g = torch.Generator().manual_seed(5)
    t = port.strat_t(torch.tensor(2.0), torch.tensor(6.0), S, torch.rand(S, generator=g))
    u_rs, u_s = torch.rand(R, S, generator=g), torch.rand(S, generator=g)
    from conftest import capture_fine_sampling
    with capture_fine_sampling() as rec:
        Cr, Cf, _ = vr.vol_render(mlp, rd.to(DEV), ro.to(DEV), num_samples=S, t=t.to(DEV), dir_norm=dn.to(DEV), hierarchical=True,
                                  _u=u_rs.to(DEV), _u_cand=u_s.to(DEV))
    loss = torch.nn.functional.mse_loss(Cr, gt.to(DEV)) + torch.nn.functional.mse_loss(Cf, gt.to(DEV))
    loss.backward()
    tables = torch.stack([e.weight.detach().cpu() for e in enc.Embedding_list]).requires_grad_()
    params = {k: v.detach().cpu() for k, v in mlp.state_dict().items()}
    near, far = torch.tensor(2.0), torch.tensor(6.0)
    # the resampler on identical input (our coarse weights): bit-identical depths
    assert torch.equal(rec["t_fine"].cpu(), port.hier_sample(rec["w"].cpu(), t, near, far, u_rs, u_s))
    # the reference's fine pass on those depths (they carry no gradient): nothing skipped, no ray excluded
    Cr_ref, Cf_ref, aux = port.vol_render(params, tables, MU, SIGMA, port.level_scales(16, 2048.0, 16), rd, ro, t, dn, 4, True,
                                          near, far, u_rs, u_s, t_fine=rec["t_fine"].cpu())
    assert torch.allclose(Cr.cpu(), Cr_ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(Cf.cpu(), Cf_ref.detach(), rtol=1e-5, atol=1e-6)
    # how many rays would have been sampled differently from the CPU's own coarse weights (last-bit differences of the
    # weights moving a cdf boundary across a draw): measured 0 of 6 on B200
    own = port.hier_sample(aux["w"].detach(), t, near, far, u_rs, u_s)
    assert (own != rec["t_fine"].cpu()).any(dim=-1).float().mean() <= 1 / 6
    (torch.nn.functional.mse_loss(Cr_ref, gt) + torch.nn.functional.mse_loss(Cf_ref, gt)).backward()
    got = torch.stack([e.weight.grad for e in enc.Embedding_list]).cpu()
    # 768-sample transmittance scans: the fp32 prefix/suffix sums of the kernel (warp scans) and of torch's cumsum
    # differ in association; measured 1.2e-4 on this gradient (the 24/128-sample fixtures hold 1e-5)
    assert float((got - tables.grad).norm() / tables.grad.norm()) < 5e-4


def test_human_config_full_step_properties():
    """configs[4] at a training batch: 1024 rays x (256 + 512) points, T = 2^22, bf16 MLP.  Properties: finite outputs;
    the table gradient has the linear structure of the scatter-add (sum over entries of level l == sum over points of
    d(feature) -- checked through a second backward with doubled upstream gradient); replays are deterministic in Cr."""
    h, enc, mlp, vr = build(2 ** 22)
    R, S = 1024, 256
    ro, rd, dn, gt = (x.to(DEV) for x in rays(R))
    torch.manual_seed(3)
    t = h.helper.strat_sampler(torch.tensor(2.0), torch.tensor(6.0), S, device=DEV)
    u_rs, u_s = torch.rand(R, S, device=DEV), torch.rand(S, device=DEV)

    def run(scale):
        for p in list(enc.parameters()) + list(mlp.parameters()):
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            Cr, Cf, _ = vr.vol_render(mlp, rd, ro, num_samples=S, t=t, dir_norm=dn, hierarchical=True, _u=u_rs, _u_cand=u_s)
            loss = scale * (torch.nn.functional.mse_loss(Cr, gt) + torch.nn.functional.mse_loss(Cf, gt))
        loss.backward()
        return Cr.detach().clone(), Cf.detach().clone(), torch.stack([e.weight.grad for e in enc.Embedding_list]).clone()

    Cr1, Cf1, g1 = run(1.0)
    Cr2, Cf2, g2 = run(2.0)
    assert bool(torch.isfinite(Cr1).all() and torch.isfinite(Cf1).all() and torch.isfinite(g1).all())
    assert torch.equal(Cr1, Cr2) and torch.equal(Cf1, Cf2)
    assert float(g1.abs().sum()) > 0
    assert float((g2 - 2 * g1).norm() / (2 * g1).norm()) < 1e-4      # linear in the upstream gradient (atomic-order noise only)
    assert Cf1.shape == (R, 3)
