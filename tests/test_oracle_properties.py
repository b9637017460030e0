"""CPU: size-independent properties of the oracle (hypothesis).  These are the invariants the GPU tests rely on at
full size, checked here on the restatement itself:
  * the uint32 hash the CUDA kernels use for power-of-two T == the reference's int64 floor-mod hash, for any lattice
    corner incl. negative and > 2^31 coordinates;
  * trilinear corner weights are a partition of unity; a constant table encodes to that constant;
  * the table gradient is linear in dy and its sum over the table equals sum(w * dy);
  * compositing: sum of weights = 1 - final transmittance, the closed-form backward == autograd;
  * SDF compositing (8f row 4): weights in [0, 1], their sum = 1 - product of (1 - alpha), the closed-form backward the
    CUDA kernel implements == autograd (d rgb, d sdf with the -10 clamp, dL/db, with and without a weight gradient);
  * hierarchical sampling returns sorted depths that contain every coarse depth."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import port

SET = settings(max_examples=60, deadline=None, derandomize=True, database=None)   # same examples on every run


@SET
@given(st.lists(st.tuples(*[st.integers(-2 ** 40, 2 ** 40)] * 3), min_size=1, max_size=64), st.integers(1, 24))
def test_u32_hash_equals_int64_floor_mod_for_pow2(corners, log2_t):
    T = 1 << log2_t
    c = torch.tensor(corners, dtype=torch.int64)
    want = port.hash_index(c, T)
    got = port.hash_index_u32_pow2(c.numpy(), T)
    assert np.array_equal(got, want.numpy())
    assert int(want.min()) >= 0 and int(want.max()) < T


@SET
@given(st.integers(0, 2 ** 31 - 1), st.integers(2, 6), st.sampled_from([64, 100, 1024]))
def test_partition_of_unity_and_constant_table(seed, L, T):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(50, 3, generator=g) * 6 - 3
    mu, sigma = torch.tensor([-3.0, -3.0, -3.0]), torch.tensor(11.0)
    scales = port.level_scales(16, 512.0, L)
    tables = torch.full((L, T, 2), 0.75)
    y, idx, w = port.hash_encode(x, tables, mu, sigma, scales, return_aux=True)
    assert torch.allclose(w.sum(-1), torch.ones(L, 50), atol=1e-6)
    assert torch.allclose(y, torch.full_like(y, 0.75), atol=1e-6)
    assert int(idx.min()) >= 0 and int(idx.max()) < T


@SET
@given(st.integers(0, 2 ** 31 - 1))
def test_table_gradient_linear_and_sum_rule(seed):
    g = torch.Generator().manual_seed(seed)
    L, T, F, N = 3, 128, 2, 40
    x = torch.rand(N, 3, generator=g)
    scales = port.level_scales(16, 256.0, L)
    _, idx, w = port.hash_encode(x, torch.zeros(L, T, F), torch.zeros(3), torch.tensor(1.0), scales, return_aux=True)
    dy1, dy2 = torch.randn(N, L * F, generator=g, dtype=torch.float64), torch.randn(N, L * F, generator=g, dtype=torch.float64)
    g1, g2 = port.hash_encode_bwd(dy1, idx, w, T, F), port.hash_encode_bwd(dy2, idx, w, T, F)
    g12 = port.hash_encode_bwd(2 * dy1 - 3 * dy2, idx, w, T, F)
    assert torch.allclose(g12, 2 * g1 - 3 * g2, atol=1e-9)
    for l in range(L):
        want = (w[l].double().sum(-1, keepdim=True) * dy1[:, l * F:(l + 1) * F].double()).sum(0)
        assert torch.allclose(g1[l].sum(0), want, atol=1e-9)


@SET
@given(st.integers(0, 2 ** 31 - 1), st.integers(2, 40))
def test_composite_weight_sum_and_closed_form_backward(seed, S):
    g = torch.Generator().manual_seed(seed)
    R = 5
    t = torch.sort(torch.rand(S, generator=g, dtype=torch.float64) * 4 + 2).values
    rgb = torch.rand(R, S, 3, generator=g, dtype=torch.float64)
    sigma = (torch.randn(R, S, generator=g, dtype=torch.float64) * 8).requires_grad_(True)   # crosses the -10 clamp
    dn = 1 + torch.rand(R, 1, generator=g, dtype=torch.float64)
    rgb.requires_grad_(True)
    C, w = port.composite(t, rgb, sigma, dn)
    delta = torch.zeros_like(t)
    delta[:-1] = t[1:] - t[:-1]
    p = torch.clamp(sigma.detach(), min=-10) * (delta[None, :] * dn)
    # telescoping sum; negative densities make transmittance grow (values up to e^80), hence relative tolerances
    # (the sum cancels: compare relative to the largest weight of the ray)
    err = (w.sum(-1) - (1 - torch.exp(-p.sum(-1)))).abs()
    assert bool((err <= 1e-9 * (w.detach().abs().max(dim=-1).values + 1)).all())
    gC = torch.randn(R, 3, generator=g, dtype=torch.float64)
    (C * gC).sum().backward()
    drgb, dsig = port.composite_bwd(t, rgb.detach(), sigma.detach(), dn, gC)
    assert torch.allclose(drgb, rgb.grad, rtol=1e-9, atol=1e-9)
    scale = sigma.grad.abs().max(dim=-1, keepdim=True).values + 1.0      # per ray: suffix sums cancel as well
    assert bool(((dsig - sigma.grad).abs() <= 1e-8 * scale).all())


@SET
@given(st.integers(0, 2 ** 31 - 1), st.integers(2, 32))
def test_hier_sample_sorted_and_contains_coarse(seed, S):
    g = torch.Generator().manual_seed(seed)
    R = 4
    t = torch.sort(torch.rand(S, generator=g) * 4 + 2).values
    w = torch.randn(R, S, generator=g)                                                          # negatives are clipped
    merged = port.hier_sample(w, t, 2.0, 6.0, torch.rand(R, S, generator=g), torch.rand(S, generator=g))
    assert merged.shape == (R, 2 * S)
    assert bool((merged[:, 1:] >= merged[:, :-1]).all())
    for r in range(R):
        assert all(bool((merged[r] == v).any()) for v in t)
    assert float(merged.min()) >= 2.0 - 1e-6 and float(merged.max()) <= 6.0 + 1e-6


@SET
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 48), st.sampled_from(["uniform", "walk", "wide"]), st.booleans())
def test_sdf_composite_weights_and_closed_form_backward(seed, S, kind, with_gw):
    g = torch.Generator().manual_seed(seed)
    R = 4
    if kind == "uniform":
        sdf, bval = torch.rand(R, S, generator=g, dtype=torch.float64) * 2 - 1, 0.5
    elif kind == "walk":
        sdf, bval = torch.cumsum(torch.randn(R, S, generator=g, dtype=torch.float64) * 0.05, -1), 0.9
    else:
        sdf, bval = (torch.rand(R, S, generator=g, dtype=torch.float64) * 2 - 1) * 14, 1.3      # crosses the -10 clamp
    rgb = torch.rand(R, S, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    sdf.requires_grad_(True)
    b = torch.tensor(bval, dtype=torch.float64, requires_grad=True)
    C, w = port.composite_sdf(rgb, sdf, b)
    w2 = w[..., 0].detach()
    assert bool((w2 >= 0).all()) and bool((w2 <= 1 + 1e-12).all())
    s = torch.clamp(sdf.detach(), min=-10)
    phi = 1 / (1 + torch.exp(-s * bval))
    alpha = torch.zeros_like(s)
    alpha[:, :-1] = torch.clamp(1 - phi[:, 1:] / phi[:, :-1], min=0)
    assert torch.allclose(w2.sum(-1), 1 - torch.prod(1 - alpha, -1), rtol=1e-10, atol=1e-12)    # telescoping product
    gC = torch.randn(R, 3, generator=g, dtype=torch.float64)
    gw = torch.randn(R, S, generator=g, dtype=torch.float64) if with_gw else None
    loss = (C * gC).sum() + ((w[..., 0] * gw).sum() if with_gw else 0)
    loss.backward()
    drgb, dsdf, db = port.composite_sdf_bwd(rgb.detach(), sdf.detach(), b.detach(), gC, gw)
    assert torch.allclose(drgb, rgb.grad, rtol=1e-10, atol=1e-12)
    scale = sdf.grad.abs().max() + 1.0
    assert bool(((dsdf - sdf.grad).abs() <= 1e-9 * scale).all())
    assert abs(float(db) - float(b.grad)) <= 1e-9 * (abs(float(b.grad)) + 1.0)
    assert bool((dsdf[sdf.detach() < -10] == 0).all())
