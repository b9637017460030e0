"""GPU tests (-m gpu): parity of the 16-bit tensor-core MLP path, per operand format, with the tolerances spelled out.

What the north star grants: outputs and gradients "within 1e-2 for the bf16 MLP".  What is achievable, measured
(profiles/r02_precision.json, scripts/measure_precision.py):

* outputs (colours, densities, loss): <= 3e-3 (bf16) / 4e-4 (fp16) of the reference's fp32 values -- asserted at 1e-2;
* gradients of a training step at the benchmarked size (C2: 4096 rays x 128 samples, summed over 524 288 points):
  table gradient 3.7e-3 (bf16) / 1.3e-3 (fp16), MLP gradients <= 5e-3 / 9e-4 norm-wise against the fp32 kernels
  (themselves pinned to the reference at 1e-5) -- asserted at 1e-2;
* PER-POINT gradients (d loss / d features of single points, or parameter gradients of ~1000 points under a random
  upstream gradient) of any 16-bit evaluation of this ReLU network differ from fp32 by percents, because roundings flip
  the ReLU mask of pre-activations near zero: the REFERENCE'S OWN autocast evaluation (tests/golden/mlp_autocast.npz,
  produced by the unmodified reference on the CPU) is 1.4-1.7 % (fp16) and 7-10 % (bf16) away from its fp32 self.
  The kernels are held to that yardstick: not worse than the reference's own 16-bit error (x1.25 + 5e-3), and closer to
  the reference's 16-bit result than that result is to fp32.
* fp16 operands need the upstream gradient scaled (fp16 underflows at ~6e-8): the reference uses GradScaler
  (train_hash2.py:156,226) -- here either the same GradScaler or MLP_3D.tc_grad_scale.
"""
import pytest
import torch

from conftest import load_golden, mlp_params
from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda"
FMTS = [("f16", torch.float16), ("bf16", torch.bfloat16)]


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def make_mlp(p):
    import human_body_reconstruction_b200 as h
    m = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=torch.ones(3), min_bound=-torch.ones(3))
    m.load_state_dict(p)
    return m.to(DEV)


@pytest.mark.parametrize("tag,fmt", FMTS)
def test_mlp_tracks_the_references_own_autocast(tag, fmt):
    g = load_golden("mlp_autocast.npz")
    m = make_mlp(mlp_params(g))
    S = int(g["S"])
    f = g["feat"].to(DEV).requires_grad_()
    out = m.field(f, g["dirs"].to(DEV), S, use_tc=fmt)
    out.backward(g["dout"].to(DEV))
    # outputs: the north star's 1e-2 against the reference's fp32 arithmetic, and as close to the reference's own
    # 16-bit result as that is to fp32
    assert rel(out, g["f32__out"]) < 1e-2
    assert rel(out, g[f"{tag}__out"]) < 1e-2
    assert rel(out, g["f32__out"]) <= 1.25 * rel(g[f"{tag}__out"], g["f32__out"]) + 1e-4
    # per-point gradients: not worse than the reference's own 16-bit evaluation
    ref_err = rel(g[f"{tag}__dfeat"], g["f32__dfeat"])
    assert rel(f.grad, g["f32__dfeat"]) <= 1.25 * ref_err + 5e-3
    assert rel(f.grad, g[f"{tag}__dfeat"]) <= ref_err          # closer to the reference's 16-bit result than that is to fp32
    for k, q in m.named_parameters():
        kk = k.replace(".", "__")
        ref_err = rel(g[f"{tag}__grad__{kk}"], g[f"f32__grad__{kk}"])
        assert rel(q.grad, g[f"f32__grad__{kk}"]) <= 1.25 * ref_err + 5e-3, k
    # absolute ceilings from the measurement (profiles/r02_precision.json): fp16 1.6 %, bf16 11 %
    ceil = 0.025 if tag == "f16" else 0.14
    assert rel(f.grad, g["f32__dfeat"]) < ceil
    assert max(rel(q.grad, g["f32__grad__" + k.replace(".", "__")]) for k, q in m.named_parameters()) < ceil


def _c2_scene(R, S, T):
    import human_body_reconstruction_b200 as h
    torch.manual_seed(0)
    mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
    sigma = ((maxb - mu) ** 2).sum().sqrt()
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=mu.to(DEV), sigma=sigma.to(DEV))
    with torch.no_grad():
        for e in enc.Embedding_list:
            e.weight.uniform_(-1.0, 1.0)                       # "trained-like" table
    enc = enc.to(DEV)
    mlp = make_mlp(port.mlp_init(seed=3))
    pe = h.PositionalEncoder(3, 4)
    near, far = torch.tensor(2.0), torch.tensor(6.0)
    vr = h.Volume_Renderer(H=800, W=800, K=torch.eye(3), near=near, far=far, device=DEV, Pos_encode=enc, Dir_encode=pe,
                           max_dim=1024, sigma_val=sigma, mu=mu)
    ro = torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.3 * torch.randn(R, 3)
    rd = torch.nn.functional.normalize(-ro + 0.8 * torch.randn(R, 3), dim=-1)
    gt = torch.rand(R, 3)
    t = port.strat_t(near, far, S, torch.rand(S))
    return vr, enc, mlp, ro.to(DEV), rd.to(DEV), gt.to(DEV), t.to(DEV)


def test_training_step_gradients_at_c2_size_within_1e_2():
    """The benchmarked step (4096 rays x 128 samples, L=16 F=2 T=2^19): colours, loss, table gradient and every MLP
    gradient of the 16-bit tensor-core path within 1e-2 (norm-wise) of the fp32 kernels."""
    R, S, T = 4096, 128, 2 ** 19
    vr, enc, mlp, ro, rd, gt, t = _c2_scene(R, S, T)

    def run(fmt, scale=1.0):
        mlp.tc_grad_scale = scale
        for q in list(enc.parameters()) + list(mlp.parameters()):
            q.grad = None
        with torch.autocast("cuda", dtype=fmt or torch.bfloat16, enabled=fmt is not None):
            Cr, _, _ = vr.vol_render(mlp, rd, ro, num_samples=S, t=t, dir_norm=1.0, hierarchical=False)
            loss = 2 * torch.nn.functional.mse_loss(Cr, gt)
        loss.backward()
        gtab = torch.stack([e.weight.grad for e in enc.Embedding_list]).clone()
        return Cr.detach().clone(), gtab, {k: q.grad.clone() for k, q in mlp.named_parameters()}, float(loss)

    base = run(None)
    report = {}
    for tag, fmt, scale in (("bf16", torch.bfloat16, 1.0), ("f16", torch.float16, 65536.0)):
        Cr, gtab, gm, loss = run(fmt, scale)
        errs = {"Cr": rel(Cr, base[0]), "loss": abs(loss - base[3]) / base[3], "table": rel(gtab, base[1]),
                "mlp": max(rel(gm[k], base[2][k]) for k in gm)}
        report[tag] = errs
        assert torch.isfinite(gtab).all()
        for k, v in errs.items():
            assert v < 1e-2, (tag, k, v)
    # measured on B200 (profiles/r02_precision.json): bf16 table 3.7e-3 / MLP 5.0e-3; fp16 1.3e-3 / 0.9e-3
    assert report["f16"]["table"] < 4e-3 and report["f16"]["mlp"] < 4e-3
    print("c2 precision:", report)


def test_fp16_operands_without_gradient_scaling_underflow_like_the_reference():
    """Documents WHY train_hash2.py wraps the step in GradScaler: fp16 pre-activation gradients of a mean-reduced loss
    underflow.  With the scale (here tc_grad_scale, there GradScaler's 65536) the gradients are right; without it the
    table gradient is far off -- and the kernel says so by being deterministic about it, not by producing inf/nan."""
    R, S, T = 512, 128, 2 ** 16
    vr, enc, mlp, ro, rd, gt, t = _c2_scene(R, S, T)

    def run(fmt, scale):
        mlp.tc_grad_scale = scale
        for q in list(enc.parameters()) + list(mlp.parameters()):
            q.grad = None
        with torch.autocast("cuda", dtype=fmt or torch.bfloat16, enabled=fmt is not None):
            Cr, _, _ = vr.vol_render(mlp, rd, ro, num_samples=S, t=t, dir_norm=1.0, hierarchical=False)
            loss = 2 * torch.nn.functional.mse_loss(Cr, gt)
        loss.backward()
        return torch.stack([e.weight.grad for e in enc.Embedding_list]).clone()

    base = run(None, 1.0)
    scaled = run(torch.float16, 65536.0)
    unscaled = run(torch.float16, 1.0)
    assert torch.isfinite(unscaled).all()
    assert rel(scaled, base) < 1e-2
    # an external GradScaler does the same job: scale the loss, unscale the gradient
    mlp.tc_grad_scale = 1.0
    for q in list(enc.parameters()) + list(mlp.parameters()):
        q.grad = None
    with torch.autocast("cuda", dtype=torch.float16):
        Cr, _, _ = vr.vol_render(mlp, rd, ro, num_samples=S, t=t, dir_norm=1.0, hierarchical=False)
        loss = 2 * torch.nn.functional.mse_loss(Cr, gt)
    (loss * 65536.0).backward()
    ext = torch.stack([e.weight.grad for e in enc.Embedding_list]) / 65536.0
    assert rel(ext, base) < 1e-2


def test_operand_format_follows_the_autocast_dtype():
    """torch.autocast(float16) -- what train_hash2.py:218 runs -- selects fp16 operands, bfloat16 selects bf16; the two
    give different (both valid) roundings of the same network."""
    from human_body_reconstruction_b200 import _lib, ops
    g = load_golden("mlp_autocast.npz")
    m = make_mlp(mlp_params(g))
    S = int(g["S"])
    f, d = g["feat"].to(DEV), g["dirs"].to(DEV)
    assert ops.tc_operand(torch.float16) == _lib.HBR_F16 and ops.tc_operand(torch.bfloat16) == _lib.HBR_BF16
    with torch.autocast("cuda", dtype=torch.float16):
        assert ops.tc_operand() == _lib.HBR_F16
        o16 = m.field(f, d, S)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert ops.tc_operand() == _lib.HBR_BF16
        obf = m.field(f, d, S)
    assert rel(o16, m.field(f, d, S, use_tc=torch.float16)) == 0.0
    assert rel(obf, m.field(f, d, S, use_tc=torch.bfloat16)) == 0.0
    assert rel(o16, g["f16__out"]) < rel(obf, g["f16__out"])          # the fp16 kernel is the one that matches fp16 autocast
    with pytest.raises(TypeError):
        ops.tc_operand(torch.float64)
