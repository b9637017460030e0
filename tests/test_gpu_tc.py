"""GPU tests (-m gpu) of the tcgen05 / TMEM MLP (csrc/mlp_tc.cu).

Three layers of evidence:
  1. the three UMMA operand modes the kernels rely on (K-major x K-major, K-major x MN-major, MN-major x MN-major
     with M=64) reproduce a bf16-input / fp32-accumulate matmul to fp32 rounding;
  2. the MLP kernels match a torch model of their own numerics (oracle.port.mlp_bf16_emulation: bf16 operands,
     fp32 accumulation) tightly -- this is the "is the kernel right" test;
  3. against the reference's fp32 arithmetic (oracle.port.mlp_forward) the outputs are within the 1e-2 the
     north star grants the 16-bit MLP.  Per-point gradients of a ReLU network evaluated in 16 bit differ from the fp32
     ones mostly through activation-mask flips at pre-activations near zero; the reference's own autocast evaluation
     shows the same (tests/golden/mlp_autocast.npz), and is the yardstick: tests/test_gpu_precision.py holds the stated
     bounds, including 1e-2 on the gradients of the benchmarked step.
Both operand formats are covered: bf16 (torch.autocast(bfloat16)) and fp16 (torch.autocast(float16), train_hash2.py:218)."""
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bf(x, fmt=torch.bfloat16):
    return x.to(fmt).float()


@pytest.mark.parametrize("mode,M,N,K", [(0, 128, 64, 32), (0, 128, 16, 64), (0, 128, 64, 48), (0, 128, 64, 64),
                                        (1, 128, 64, 16), (1, 128, 48, 64), (1, 128, 32, 64), (1, 128, 64, 64),
                                        (2, 64, 64, 128), (2, 64, 48, 128), (2, 64, 32, 128), (2, 64, 16, 128)])
@pytest.mark.parametrize("fmt", [torch.bfloat16, torch.float16])
def test_umma_operand_modes(mode, M, N, K, fmt):
    from human_body_reconstruction_b200 import ops
    torch.manual_seed(mode * 1000 + N + K)
    if mode == 0:
        A, B = torch.randn(M, K), torch.randn(N, K)
        ref = bf(A, fmt) @ bf(B, fmt).T
    elif mode == 1:
        A, B = torch.randn(M, K), torch.randn(K, N)
        ref = bf(A, fmt) @ bf(B, fmt)
    else:
        A, B = torch.randn(K, 64), torch.randn(K, N)
        ref = bf(A, fmt).T @ bf(B, fmt)
    D = ops.debug_umma(mode, A.to(DEV), B.to(DEV), M, N, K, operand=ops.tc_operand(fmt))
    assert rel(D, ref) < 1e-6


def make(seed=5):
    import human_body_reconstruction_b200 as h
    p = port.mlp_init(seed=seed)
    m = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=torch.ones(3), min_bound=-torch.ones(3))
    m.load_state_dict(p)
    return p, m.to(DEV)


@pytest.mark.parametrize("R,S", [(1, 1), (3, 100), (40, 24), (512, 128), (129, 7)])
@pytest.mark.parametrize("fmt", [torch.bfloat16, torch.float16])
def test_mlp_tc_matches_numerics_model_and_reference(R, S, fmt):
    torch.manual_seed(R * 1000 + S)
    p, m = make()
    feat = torch.randn(R * S, 32) * 0.5
    dirs = port.dir_encode(torch.nn.functional.normalize(torch.randn(R, 3), dim=-1), 4)
    drep = dirs[:, None, :].repeat(1, S, 1).reshape(R * S, -1)
    dout = torch.randn(R * S, 4)
    emu_out, emu_dfeat, emu_g, _ = port.mlp_tc_emulation(p, feat, drep, dout, fmt)
    pr = {k: v.clone().requires_grad_() for k, v in p.items()}
    fr = feat.clone().requires_grad_()
    ref = port.mlp_forward(pr, fr, drep)
    ref.backward(dout)

    f = feat.to(DEV).requires_grad_()
    out = m.field(f, dirs.to(DEV), S, use_tc=fmt)
    out.backward(dout.to(DEV))
    # (2) kernel vs its numerics model
    # a last-bit difference in an fp32 accumulator flips the 16-bit rounding of that activation (2^-9 / 2^-12 relative for
    # the element); a handful of flips per row is the expected residual between two correct evaluations
    assert rel(out, emu_out) < 5e-4
    # ReLU-mask / rounding flips of single activations dominate the residual between two correct evaluations;
    # on a few hundred points one flip is already a percent of a bias gradient
    gtol = 1e-2 if R * S >= 4096 else 3e-2
    assert rel(f.grad, emu_dfeat) < gtol
    for k, q in m.named_parameters():
        assert rel(q.grad, emu_g[k]) < gtol, k
    # (3) kernel vs the reference's fp32 arithmetic: outputs within the north star's 1e-2
    assert rel(out, ref) < 1e-2
    if R * S >= 65536:
        # per-point gradients against fp32: the yardstick is what the reference's own torch expressions give under
        # torch.autocast of the same dtype (oracle.port.mlp_forward == the reference's ops, pinned on the CPU by
        # tests/test_oracle_golden.py::test_mlp_autocast_golden); tests/test_gpu_precision.py states the bounds
        pa = {k: v.clone().requires_grad_() for k, v in p.items()}
        fa = feat.clone().requires_grad_()
        with torch.autocast("cpu", dtype=fmt):
            oa = port.mlp_forward(pa, fa, drep)
        oa.float().backward(dout)
        assert rel(f.grad, fr.grad) <= 1.25 * rel(fa.grad, fr.grad) + 5e-3
        for k, q in m.named_parameters():
            assert rel(q.grad, pr[k].grad) <= 1.25 * rel(pa[k].grad, pr[k].grad) + 5e-3, k


def test_mlp_tc_under_autocast_is_selected_and_accumulates():
    """Under torch.autocast the module picks the tensor-core path; ddirs and repeated backward calls accumulate."""
    torch.manual_seed(1)
    p, m = make()
    R, S = 16, 32
    feat = (torch.randn(R * S, 32) * 0.5).to(DEV).requires_grad_()
    dirs = port.dir_encode(torch.nn.functional.normalize(torch.randn(R, 3), dim=-1), 4).to(DEV).requires_grad_()
    from human_body_reconstruction_b200 import _lib
    _lib.STATS.reset()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m.field(feat, dirs, S)
    assert _lib.STATS.calls.get("hbr_mlp_fwd_tc", 0) == 1 and "hbr_mlp_fwd_f32" not in _lib.STATS.calls
    dout = torch.randn(R * S, 4, device=DEV)
    out.backward(dout, retain_graph=True)
    g1 = {k: q.grad.clone() for k, q in m.named_parameters()}
    dd1 = dirs.grad.clone()
    out.backward(dout)
    for k, q in m.named_parameters():
        assert rel(q.grad, 2 * g1[k]) < 1e-5
    assert rel(dirs.grad, 2 * dd1) < 1e-5
    drep = dirs.detach().cpu()[:, None, :].repeat(1, S, 1).reshape(R * S, -1)
    _, _, _, emu_dd = port.mlp_tc_emulation(p, feat.detach().cpu(), drep, dout.cpu(), torch.bfloat16)
    assert rel(dd1, emu_dd.reshape(R, S, -1).sum(1)) < 5e-3


@pytest.mark.parametrize("fmt,scale,gtol", [(torch.bfloat16, 1.0, 0.05), (torch.float16, 65536.0, 0.03)])
def test_vol_render_16bit_close_to_reference(fmt, scale, gtol):
    """Whole coarse render under autocast against the reference's fp32 fixture (64 rays x 24 samples): colours and loss
    within 1e-2.  The table gradient of these 1 536 points is in the per-point regime (see tests/test_gpu_precision.py):
    measured on B200 2.8 % (bf16) / 1.7 % (fp16, gradient scale 2^16); the bound is that with margin.  At the benchmarked size the same quantity is within 1e-2
    (test_gpu_precision.py::test_training_step_gradients_at_c2_size_within_1e_2)."""
    from conftest import load_golden, mlp_params
    from test_gpu_parity import build_renderer
    g = load_golden("volrender.npz")
    vr, enc, mlp = build_renderer(g)
    mlp.tc_grad_scale = scale
    S = 24
    t = port.strat_t(g["near"], g["far"], S, g["coarse__u_t"]).to(DEV)
    with torch.autocast("cuda", dtype=fmt):
        Cr, Cf, _ = vr.vol_render(mlp, g["rays_d"].to(DEV), g["rays_o"].to(DEV), num_samples=S, t=t,
                                  dir_norm=g["dir_norm"].to(DEV), hierarchical=False)
        loss = torch.nn.functional.mse_loss(Cr, g["gt"].to(DEV)) + torch.nn.functional.mse_loss(Cf, g["gt"].to(DEV))
    assert rel(Cr, g["coarse__Cr"]) < 1e-2
    assert abs(float(loss.detach()) - float(g["coarse__loss"])) < 1e-2 * float(g["coarse__loss"])
    loss.backward()
    grad = torch.stack([e.weight.grad for e in enc.Embedding_list])
    err = rel(grad, g["coarse__dtables"])
    print(f"volrender fixture, {fmt}: table-gradient error {err:.4f}")
    assert torch.isfinite(grad).all() and err < gtol


@pytest.mark.parametrize("R,S,T", [(96, 128, 2 ** 19), (33, 40, 2 ** 14), (1, 7, 2 ** 10)])
def test_fused_field_matches_separate_kernels(R, S, T):
    """hbr_field_fwd_tc / hbr_field_bwd_tc (encoder + MLP in one kernel) against hbr_hash_encode_* + hbr_mlp_*_tc:
    same bf16 features enter the same GEMM chain, so outputs agree to fp32 rounding and gradients to atomic-order noise."""
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200 import ops
    from human_body_reconstruction_b200.vol_renderer import _FieldFn
    torch.manual_seed(R + S)
    mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
    sigma = ((maxb - mu) ** 2).sum().sqrt()
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=mu.to(DEV), sigma=sigma.to(DEV))
    with torch.no_grad():
        for e in enc.Embedding_list:
            e.weight.mul_(5e3)
    enc = enc.to(DEV)
    p, m = make()
    pts = (mu + (maxb - mu) * (0.1 + 0.45 * torch.rand(R * S, 3))).to(DEV)
    # consecutive samples along rays share cells on the coarse levels: exercise the run merging of the scatter
    pts = pts.view(R, S, 3)
    pts[:, 1:, :] = pts[:, :1, :] + torch.linspace(0, 1, S - 1, device=DEV)[None, :, None] * 0.05 if S > 1 else pts[:, 1:, :]
    pts = pts.reshape(-1, 3).contiguous()
    dirs = port.dir_encode(torch.nn.functional.normalize(torch.randn(R, 3), dim=-1), 4).to(DEV)
    dout = torch.randn(R * S, 4, device=DEV)

    def grads():
        gt = torch.stack([e.weight.grad for e in enc.Embedding_list]).clone()
        gm = {k: q.grad.clone() for k, q in m.named_parameters()}
        for q in list(enc.parameters()) + list(m.parameters()):
            q.grad = None
        return gt, gm

    out_a = m.field(enc(pts), dirs, S, use_tc=True)
    out_a.backward(dout)
    gt_a, gm_a = grads()
    out_b = _FieldFn.apply(pts, dirs, S, enc, m, *[e.weight for e in enc.Embedding_list], *m._ordered())
    out_b.backward(dout)
    gt_b, gm_b = grads()
    assert rel(out_b, out_a) < 1e-6
    assert rel(gt_b, gt_a) < 1e-5
    for k in gm_a:
        assert rel(gm_b[k], gm_a[k]) < 1e-4, k


def test_vol_render_fused_option_matches_default():
    """Volume_Renderer.fuse_field=True routes the autocast pass through the fused kernels; colours and gradients agree."""
    from conftest import load_golden
    from test_gpu_parity import build_renderer
    g = load_golden("volrender.npz")
    res = []
    for fuse in (False, True):
        vr, enc, mlp = build_renderer(g)
        vr.fuse_field = fuse
        S = 24
        t = port.strat_t(g["near"], g["far"], S, g["coarse__u_t"]).to(DEV)
        from human_body_reconstruction_b200 import _lib
        _lib.STATS.reset()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            Cr, Cf, _ = vr.vol_render(mlp, g["rays_d"].to(DEV), g["rays_o"].to(DEV), num_samples=S, t=t,
                                      dir_norm=g["dir_norm"].to(DEV), hierarchical=False)
            loss = torch.nn.functional.mse_loss(Cr, g["gt"].to(DEV))
        loss.backward()
        assert ("hbr_field_fwd_tc" in _lib.STATS.calls) == fuse and ("hbr_field_bwd_tc" in _lib.STATS.calls) == fuse
        res.append((Cr.detach(), torch.stack([e.weight.grad for e in enc.Embedding_list])))
    assert rel(res[1][0], res[0][0]) < 1e-5
    assert rel(res[1][1], res[0][1]) < 1e-4


@pytest.mark.parametrize("fmt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("hier", [False, True])
def test_chained_field_path_matches_module_path(fmt, hier):
    """Volume_Renderer.chain_field (one autograd node: positions formed inside the hash kernels, 16-bit features handed to
    the MLP kernels) against the module-by-module path (HashEncoder.forward -> MLP_3D.field): the MLP sees bit-identical
    inputs, so colours are bit-identical and gradients agree to atomic-order noise."""
    from conftest import load_golden
    from test_gpu_parity import build_renderer
    from human_body_reconstruction_b200 import _lib
    g = load_golden("volrender.npz")
    res = []
    for chain in (False, True):
        vr, enc, mlp = build_renderer(g)
        vr.chain_field = chain
        mlp.tc_grad_scale = 4096.0 if fmt == torch.float16 else 1.0
        S = 24
        t = port.strat_t(g["near"], g["far"], S, g["coarse__u_t"]).to(DEV)
        _lib.STATS.reset()
        with torch.autocast("cuda", dtype=fmt):
            Cr, Cf, _ = vr.vol_render(mlp, g["rays_d"].to(DEV), g["rays_o"].to(DEV), num_samples=S, t=t,
                                      dir_norm=g["dir_norm"].to(DEV), hierarchical=hier, _u=g["hier__u_rs"][:, :S].to(DEV),
                                      _u_cand=g["hier__u_s"][:S].to(DEV))
            loss = torch.nn.functional.mse_loss(Cr, g["gt"].to(DEV)) + torch.nn.functional.mse_loss(Cf, g["gt"].to(DEV))
        loss.backward()
        assert ("hbr_hash_encode_fwd_rays" in _lib.STATS.calls) == chain
        assert ("hbr_hash_encode_bwd_rays" in _lib.STATS.calls or "hbr_field_bwd_rays_tc" in _lib.STATS.calls) == chain
        assert ("hbr_ray_points" in _lib.STATS.calls) == (not chain or hier)     # the chained path needs no position tensor
        res.append((Cr.detach(), Cf.detach(), torch.stack([e.weight.grad for e in enc.Embedding_list]),
                    {k: q.grad.clone() for k, q in mlp.named_parameters()}))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert rel(res[1][2], res[0][2]) < 1e-5
    for k in res[0][3]:
        assert rel(res[1][3][k], res[0][3][k]) < 1e-5, k


def test_hash_encode_from_rays_matches_positions_path():
    """hbr_hash_encode_fwd_rays / _bwd_rays == hbr_ray_points followed by hbr_hash_encode_fwd / _bwd, bit for bit in fp32;
    the 16-bit outputs are the round-to-nearest of the fp32 ones."""
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200 import ops, _lib
    torch.manual_seed(2)
    mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
    sigma = ((maxb - mu) ** 2).sum().sqrt()
    for T, R, S, per_ray in ((2 ** 14, 37, 40, False), (1000, 5, 128, True), (2 ** 12, 129, 7, True)):
        enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=mu.to(DEV), sigma=sigma.to(DEV))
        with torch.no_grad():
            for e in enc.Embedding_list:
                e.weight.mul_(5e3)
        enc = enc.to(DEV)
        ro = (torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.3 * torch.randn(R, 3)).to(DEV)
        rd = torch.nn.functional.normalize(-ro.cpu() + 0.8 * torch.randn(R, 3), dim=-1).to(DEV)
        t = (2 + 4 * torch.rand(R, S)).sort(-1).values.to(DEV) if per_ray else torch.linspace(2, 6, S, device=DEV)
        geom, table = enc._geom(), enc._flat_table()
        pts = ops.ray_points(ro, rd, t).view(-1, 3)
        y = ops.hash_encode_fwd(pts, table, geom)
        y_r = ops.hash_encode_fwd_rays(ro, rd, t, table, geom, _lib.HBR_F32)
        assert torch.equal(y, y_r)
        assert torch.equal(ops.hash_encode_fwd_rays(ro, rd, t, table, geom, _lib.HBR_BF16), y.bfloat16())
        assert torch.equal(ops.hash_encode_fwd_rays(ro, rd, t, table, geom, _lib.HBR_F16), y.half())
        dy = torch.randn_like(y)
        g0, g1 = torch.zeros_like(table), torch.zeros_like(table)
        ops.hash_encode_bwd(pts, dy, geom, g0)
        ops.hash_encode_bwd_rays(ro, rd, t, dy, geom, g1)
        assert rel(g1, g0) < 1e-6


@pytest.mark.parametrize("R,S,T,per_ray", [(1, 1, 2 ** 10, False), (3, 100, 2 ** 14, True), (129, 7, 2 ** 12, True),
                                           (700, 128, 2 ** 19, False), (4096, 32, 2 ** 19, True)])
@pytest.mark.parametrize("fmt", [torch.bfloat16, torch.float16])
def test_field_bwd_rays_matches_two_kernels(R, S, T, per_ray, fmt):
    """hbr_field_bwd_rays_tc (the scatter-add on dedicated warps of the MLP backward kernel) against hbr_mlp_bwd_tc followed by
    hbr_hash_encode_bwd_rays: the same d(features) reach the same run-merged reductions, so the MLP gradients are
    bit-identical up to the per-CTA row sum and the table gradient agrees to atomic-order noise.  Covers a single partial
    tile, several tiles per slot (700 * 128 / 128 = 700 tiles > 296 slots) and per-ray depths."""
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200 import ops, _lib
    torch.manual_seed(R * 7 + S)
    operand = _lib.HBR_BF16 if fmt == torch.bfloat16 else _lib.HBR_F16
    scale = 1.0 if fmt == torch.bfloat16 else 1024.0
    mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
    sigma = ((maxb - mu) ** 2).sum().sqrt()
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=mu.to(DEV), sigma=sigma.to(DEV))
    with torch.no_grad():
        for e in enc.Embedding_list:
            e.weight.mul_(5e3)
    enc = enc.to(DEV)
    p, m = make()
    flat, dims, geom, table = m._flat_params(), m._dims(), enc._geom(), enc._flat_table()
    assert ops.field_scatter_supported(geom, dims)
    ro = (torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.3 * torch.randn(R, 3)).to(DEV)
    rd = torch.nn.functional.normalize(-ro.cpu() + 0.8 * torch.randn(R, 3), dim=-1).to(DEV)
    t = (2 + 4 * torch.rand(R, S)).sort(-1).values.to(DEV) if per_ray else torch.linspace(2, 6, S, device=DEV)
    dirs = port.dir_encode(torch.nn.functional.normalize(torch.randn(R, 3), dim=-1), 4).to(DEV)
    feat16 = ops.hash_encode_fwd_rays(ro, rd, t, table, geom, operand)
    out, _ = ops.mlp_fwd_tc(feat16, dirs, S, flat, dims, operand=operand)
    dout = torch.randn(R * S, 4, device=DEV) / (R * S)
    g0, g1 = torch.zeros_like(table), torch.zeros_like(table)
    dp0, dp1 = torch.zeros_like(flat), torch.zeros_like(flat)
    dfeat, dd0 = ops.mlp_bwd_tc(feat16, dirs, S, flat, dims, out, dout, True, True, dp0, operand=operand, grad_scale=scale)
    ops.hash_encode_bwd_rays(ro, rd, t, dfeat, geom, g0)
    dd1 = ops.field_bwd_rays_tc(feat16, ro, rd, t, geom, dirs, flat, dims, out, dout, g1, True, dp1, operand=operand,
                                grad_scale=scale)
    torch.cuda.synchronize()
    assert torch.isfinite(g1).all() and g0.abs().sum() > 0
    assert rel(g1, g0) < 1e-6
    assert rel(dp1, dp0) < 1e-6 and rel(dd1, dd0) < 1e-6
    # accumulation contract: a second call adds on top
    ops.field_bwd_rays_tc(feat16, ro, rd, t, geom, dirs, flat, dims, out, dout, g1, False, dp1, operand=operand, grad_scale=scale)
    assert rel(g1, 2 * g0) < 1e-6 and rel(dp1, 2 * dp0) < 1e-6


@pytest.mark.parametrize("R,S,T,per_ray", [(1, 2, 2 ** 10, False), (3, 100, 2 ** 14, True), (130, 7, 2 ** 12, True),
                                           (700, 128, 2 ** 19, False), (4096, 32, 2 ** 19, True)])
@pytest.mark.parametrize("fmt", [torch.bfloat16, torch.float16])
def test_level_major_backward_pair_matches_row_major(R, S, T, per_ray, fmt):
    """The multi-GPU form of the step's backward: hbr_mlp_bwd_tc writing d(features) LEVEL-MAJOR (16, N, 2)
    (dfeat_stride = HBR_DFEAT_LEVEL_MAJOR) followed by the level-major scatter-add hbr_hash_encode_bwd_rays_lm, against the
    row-major pair.  The level-major tensor is the exact transpose of the row-major one (same values, bit for bit), the MLP
    gradients do not depend on the layout, and the table gradient agrees to atomic-order noise -- with and without level
    chunks / completion counters, and through the explicit-positions entry point.  Covers partial tiles and partial CTAs
    of the 1024-point level-major grid."""
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200 import ops, _lib
    torch.manual_seed(R * 5 + S)
    operand = _lib.HBR_BF16 if fmt == torch.bfloat16 else _lib.HBR_F16
    scale = 1.0 if fmt == torch.bfloat16 else 1024.0
    mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
    sigma = ((maxb - mu) ** 2).sum().sqrt()
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=mu.to(DEV), sigma=sigma.to(DEV))
    with torch.no_grad():
        for e in enc.Embedding_list:
            e.weight.mul_(5e3)
    enc = enc.to(DEV)
    p, m = make()
    flat, dims, geom, table = m._flat_params(), m._dims(), enc._geom(), enc._flat_table()
    N = R * S
    assert ops.mlp_level_major_ok(geom, dims, N)
    ro = (torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.3 * torch.randn(R, 3)).to(DEV)
    rd = torch.nn.functional.normalize(-ro.cpu() + 0.8 * torch.randn(R, 3), dim=-1).to(DEV)
    t = (2 + 4 * torch.rand(R, S)).sort(-1).values.to(DEV) if per_ray else torch.linspace(2, 6, S, device=DEV)
    dirs = port.dir_encode(torch.nn.functional.normalize(torch.randn(R, 3), dim=-1), 4).to(DEV)
    feat16 = ops.hash_encode_fwd_rays(ro, rd, t, table, geom, operand)
    out, _ = ops.mlp_fwd_tc(feat16, dirs, S, flat, dims, operand=operand)
    dout = torch.randn(N, 4, device=DEV) / N
    dp0, dp1 = torch.zeros_like(flat), torch.zeros_like(flat)
    dfeat, dd0 = ops.mlp_bwd_tc(feat16, dirs, S, flat, dims, out, dout, True, True, dp0, operand=operand, grad_scale=scale)
    dfeat_lm, dd1 = ops.mlp_bwd_tc(feat16, dirs, S, flat, dims, out, dout, True, True, dp1, operand=operand, grad_scale=scale,
                                   level_major=True)
    assert dfeat_lm.shape == (16, N, 2)
    assert torch.equal(dfeat_lm, dfeat.view(N, 16, 2).permute(1, 0, 2).contiguous())
    assert rel(dp1, dp0) < 1e-6 and rel(dd1, dd0) < 1e-6          # (accumulated with atomics: order noise only)
    g0, g1, g2, g3 = (torch.zeros_like(table) for _ in range(4))
    ops.hash_encode_bwd_rays(ro, rd, t, dfeat, geom, g0)
    ops.hash_encode_bwd_rays_lm(ro, rd, t, dfeat_lm, geom, g1)
    chunks = [(0, 3), (3, 4), (4, 10), (10, 16)]
    done = torch.zeros(8, dtype=torch.int32, device=DEV)
    ops.hash_encode_bwd_rays_lm(ro, rd, t, dfeat_lm, geom, g2, chunks, done)
    tt = t if per_ray else t[None, :].expand(R, S)
    pts = (ro[:, None, :] + rd[:, None, :] * tt[:, :, None]).reshape(-1, 3).contiguous()
    ops.hash_encode_bwd_lm(pts, dfeat_lm, geom, g3, chunks[:2] + [(4, 16)], done[4:])
    torch.cuda.synchronize()
    assert g0.abs().sum() > 0
    assert rel(g1, g0) < 1e-6 and rel(g2, g0) < 1e-6 and rel(g3, g0) < 1e-6
    ctas = ops.hash_bwd_lm_ctas(N)
    assert ctas == -(-N // 1024) and done.tolist() == [ctas] * 4 + [ctas] * 3 + [0]


@pytest.mark.parametrize("R,S,T,per_ray", [(1, 1, 2 ** 10, False), (3, 100, 2 ** 14, True), (129, 7, 2 ** 12, True),
                                           (700, 128, 2 ** 19, False), (4096, 32, 2 ** 19, True)])
@pytest.mark.parametrize("fmt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("image", [False, True])
def test_field_fwd_rays_matches_two_kernels(R, S, T, per_ray, fmt, image):
    """hbr_field_fwd_rays_tc (the hash-grid gather on dedicated warps of the MLP forward kernel, feeding its tile groups
    through shared memory) against hbr_hash_encode_fwd_rays followed by hbr_mlp_fwd_tc: the same arithmetic on the same
    operands, so the 16-bit features AND the (rgb, sigma) outputs are bit-identical -- whether the kernel converts the
    parameters itself or copies the prepared operand image.  Covers a single partial tile, several tiles per tile group
    (700 * 128 / 128 = 700 tiles > 296 slots), per-ray depths, and a ragged last tile."""
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200 import ops, _lib
    torch.manual_seed(R * 3 + S)
    operand = _lib.HBR_BF16 if fmt == torch.bfloat16 else _lib.HBR_F16
    mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
    sigma = ((maxb - mu) ** 2).sum().sqrt()
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=mu.to(DEV), sigma=sigma.to(DEV))
    with torch.no_grad():
        for e in enc.Embedding_list:
            e.weight.mul_(5e3)
    enc = enc.to(DEV)
    p, m = make()
    flat, dims, geom, table = m._flat_params(), m._dims(), enc._geom(), enc._flat_table()
    assert ops.field_scatter_supported(geom, dims)
    ro = (torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.3 * torch.randn(R, 3)).to(DEV)
    rd = torch.nn.functional.normalize(-ro.cpu() + 0.8 * torch.randn(R, 3), dim=-1).to(DEV)
    t = (2 + 4 * torch.rand(R, S)).sort(-1).values.to(DEV) if per_ray else torch.linspace(2, 6, S, device=DEV)
    dirs = port.dir_encode(torch.nn.functional.normalize(torch.randn(R, 3), dim=-1), 4).to(DEV)
    feat16 = ops.hash_encode_fwd_rays(ro, rd, t, table, geom, operand)
    out, _ = ops.mlp_fwd_tc(feat16, dirs, S, flat, dims, operand=operand)
    if image:
        ops.mlp_tc_prepare(flat, dims, operand)                 # the fused kernel copies the image
    else:
        ops._tc_image.clear()                                   # ... or converts the fp32 parameters itself
    out1, feat1 = ops.field_fwd_rays_tc(ro, rd, t, table, geom, dirs, flat, dims, operand)
    torch.cuda.synchronize()
    assert feat1.dtype == feat16.dtype and torch.equal(feat1.view(torch.int16), feat16.view(torch.int16))
    assert torch.isfinite(out1).all() and torch.equal(out1, out)
