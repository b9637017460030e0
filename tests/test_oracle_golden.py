"""CPU: oracle/port.py against the committed reference outputs (tests/golden, made by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, mlp_params
from oracle import port


@pytest.mark.parametrize("tag", ["pow2", "npow2"])
def test_hash_indices_bit_exact(tag):
    g = load_golden(f"hash_{tag}.npz")
    L, T, F = g["tables"].shape
    scales = port.level_scales(int(g["n_min"]), float(g["n_max"]), L)
    assert torch.equal(scales, g["scales"])
    y, idx, w = port.hash_encode(g["x"], g["tables"], g["mu"], g["sigma"], scales, return_aux=True)
    assert torch.equal(idx, g["idx"])
    assert idx.min() >= 0 and idx.max() < T
    assert torch.equal(y, g["y"])
    assert torch.equal(port.hash_encode(g["x"].half(), g["tables"], g["mu"], g["sigma"], scales), g["y_from_f16"])
    if tag == "pow2":   # the uint32 form used on the GPU is the same function for power-of-two T
        for l in range(L):
            x0, _ = port.hash_cells(g["x"], g["mu"], g["sigma"], scales[l])
            c = (x0[:, None, :] + port._CORNER_BITS[None]).numpy()
            assert np.array_equal(port.hash_index_u32_pow2(c, T), idx[l].numpy())
        assert (port.hash_cells(g["x"], g["mu"], g["sigma"], scales[-1])[0] < 0).any()   # negative cells covered


@pytest.mark.parametrize("tag", ["pow2", "npow2"])
def test_hash_backward(tag):
    g = load_golden(f"hash_{tag}.npz")
    L, T, F = g["tables"].shape
    _, idx, w = port.hash_encode(g["x"], g["tables"], g["mu"], g["sigma"], g["scales"], return_aux=True)
    d = port.hash_encode_bwd(g["dy"], idx, w, T, F)
    ref = g["dtables"].double()
    assert (d - ref).norm() / ref.norm() < 1e-6


def test_dir_encode():
    g = load_golden("dir.npz")
    assert torch.equal(port.dir_encode(g["d"], 4), g["enc"])
    assert torch.equal(port.dir_encode(g["d"].half(), 4).float(), g["enc_from_f16"])


def test_mlp_forward_backward():
    g = load_golden("mlp.npz")
    p = {k: v.clone().requires_grad_() for k, v in mlp_params(g).items()}
    assert sum(v.numel() for v in p.values()) == 14227
    feat = g["feat"].clone().requires_grad_()
    dirs = g["dirs"].clone().requires_grad_()
    out = port.mlp_forward(p, feat, dirs)
    assert torch.equal(out, g["out"])
    assert torch.equal(port.mlp_forward(p, feat, None), g["density_only"])
    out.backward(g["dout"])
    assert torch.allclose(feat.grad, g["dfeat"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(dirs.grad, g["ddirs"], rtol=1e-6, atol=1e-7)
    for k, v in p.items():
        assert torch.allclose(v.grad, g["grad__" + k.replace(".", "__")], rtol=1e-5, atol=1e-6), k


@pytest.mark.parametrize("name", ["composite.npz", "composite_perray.npz"])
def test_composite(name):
    g = load_golden(name)
    C, w = port.composite(g["t"], g["rgb"], g["sigma"], g["dir_norm"])
    assert torch.equal(C, g["C"]) and torch.equal(w, g["w"])
    drgb, dsig = port.composite_bwd(g["t"], g["rgb"], g["sigma"], g["dir_norm"], g["gC"])
    assert torch.allclose(drgb, g["drgb"], rtol=1e-6, atol=1e-7)
    assert (dsig - g["dsigma"]).norm() / g["dsigma"].norm() < 1e-5
    # fp64 closed form agrees with the reference's fp32 autograd to fp32 rounding
    d64 = port.composite_bwd(*(g[k].double() for k in ("t", "rgb", "sigma", "dir_norm", "gC")))[1]
    assert (g["dsigma"].double() - d64).norm() / d64.norm() < 1e-5


def test_hier_sample():
    g = load_golden("hier.npz")
    tf = port.hier_sample(g["w"], g["t"], g["near"], g["far"], g["u_rs"], g["u_s"])
    assert torch.equal(tf, g["t_fine"])
    assert torch.equal(port.ray_points(g["rays_o"], g["rays_d"], tf), g["rays_fine"])
    assert (tf[:, 1:] >= tf[:, :-1]).all()


@pytest.mark.parametrize("tag", ["coarse", "hier"])
def test_vol_render(tag):
    g = load_golden("volrender.npz")
    p = {k: v.clone().requires_grad_() for k, v in mlp_params(g, "mlp__").items()}
    tables = g["tables"].clone().requires_grad_()
    S = g[f"{tag}__u_t"].shape[0]
    t = port.strat_t(g["near"], g["far"], S, g[f"{tag}__u_t"])
    hier = tag == "hier"
    Cr, Cf, _ = port.vol_render(p, tables, g["mu"], g["sigma"], g["scales"], g["rays_d"], g["rays_o"], t, g["dir_norm"],
                                4, hier, g["near"], g["far"], g.get("hier__u_rs"), g.get("hier__u_s"))
    assert torch.equal(Cr.detach(), g[f"{tag}__Cr"]) and torch.equal(Cf.detach(), g[f"{tag}__Cf"])
    loss = torch.nn.functional.mse_loss(Cr, g["gt"]) + torch.nn.functional.mse_loss(Cf, g["gt"])
    loss.backward()
    ref = g[f"{tag}__dtables"]
    assert (tables.grad - ref).norm() / ref.norm() < 1e-5
    for k, v in p.items():
        r = g[f"{tag}__grad__" + k.replace(".", "__")]
        assert (v.grad - r).norm() <= 1e-5 * r.norm() + 1e-9, k


def test_grid_query():
    g = load_golden("grid.npz")
    res = int(g["res"])
    pts = port.grid_points(g["min_bound"].numpy(), g["max_bound"].numpy(), res)
    assert torch.equal(pts.float(), g["grid_f16"])
    out = port.grid_query(mlp_params(g, "mlp__"), g["tables"], g["mu"], g["sigma"], g["scales"], pts, batch=500)
    # batch size changes the CPU GEMM blocking -> last-ulp differences only
    assert torch.allclose(out.reshape(res, res, res, 4), g["out"], rtol=1e-5, atol=1e-6)
    out1 = port.grid_query(mlp_params(g, "mlp__"), g["tables"], g["mu"], g["sigma"], g["scales"], pts)
    assert torch.equal(out1.reshape(res, res, res, 4), g["out"])


def test_rays_and_bbox():
    g = load_golden("rays.npz")
    o, d, n = port.get_od(int(g["H"]), int(g["W"]), g["K"], g["c2w"])
    assert torch.equal(o, g["rays_o"]) and torch.equal(d, g["rays_d"]) and torch.equal(n, g["dir_norm"])
    mx, mn = port.bounding_box(g["c2w"], g["K"], 2.0, 6.0)
    assert torch.allclose(mx, g["max_bound"]) and torch.allclose(mn, g["min_bound"])


def test_mc_crossing_edges_small():
    rng = np.random.default_rng(0)
    d = rng.normal(30, 5, size=(7, 6, 5)).astype(np.float32)
    n = port.mc_crossing_edges(d, 30.0)
    # brute force
    ins = d < 30.0
    b = 0
    for i in range(7):
        for j in range(6):
            for k in range(5):
                if i + 1 < 7 and ins[i, j, k] != ins[i + 1, j, k]: b += 1
                if j + 1 < 6 and ins[i, j, k] != ins[i, j + 1, k]: b += 1
                if k + 1 < 5 and ins[i, j, k] != ins[i, j, k + 1]: b += 1
    assert n == b
    ci = port.mc_case_index(d, 30.0)
    assert ci.shape == (6, 5, 4)
    assert ci[0, 0, 0] == sum(int(ins[v & 1, (v >> 1) & 1, (v >> 2) & 1]) << v for v in range(8))


def test_sdf_mode_end_to_end():
    """SDF mode (8f row 4): the oracle's restatement against the reference's vol_render(use_sdf=True) fixture -- colours,
    eikonal norms, forward_sdf values, the loss of train_hash2.py:221-224 and its gradients w.r.t. the VarModel
    sharpness, the hash tables and the MLP."""
    g = load_golden("sdf.npz")
    p = {k: v.clone().requires_grad_(True) for k, v in mlp_params(g, "mlp__").items()}
    tables = g["tables"].clone().requires_grad_(True)
    b = g["b"].clone().requires_grad_(True)
    Cr, wts, norm = port.vol_render_sdf(p, tables, g["mu"], g["sigma"], g["scales"], g["rays_d"], g["rays_o"], g["t"], b,
                                        g["min_bound"], g["max_bound"])
    assert torch.allclose(Cr, g["Cr"], rtol=1e-5, atol=1e-8)
    assert torch.allclose(norm, g["norm"], rtol=1e-4, atol=1e-6)
    pts = port.ray_points(g["rays_o"], g["rays_d"], g["t"]).reshape(-1, 3)
    sdf = port.mlp_forward_sdf(p, port.hash_encode(pts, tables, g["mu"], g["sigma"], g["scales"]), None)
    assert torch.allclose(sdf, g["sdf"], rtol=1e-6, atol=1e-8)
    loss = 2 * torch.nn.functional.mse_loss(Cr, g["gt"]) + 0.1 * torch.mean((norm - 1) ** 2)
    assert torch.allclose(loss, g["loss"], rtol=1e-6)
    loss.backward()
    assert torch.allclose(b.grad, g["grad_b"], rtol=1e-4, atol=1e-9)
    rel = lambda a, w: float((a - w).norm() / w.norm())
    assert rel(tables.grad, g["dtables"]) < 1e-5
    for k, v in p.items():
        assert rel(v.grad, g["grad__" + k.replace(".", "__")]) < 1e-5, k


@pytest.mark.parametrize("tag,fmt", [("f32", None), ("f16", torch.float16), ("bf16", torch.bfloat16)])
def test_mlp_autocast_golden(tag, fmt):
    """oracle.port.mlp_forward is the reference's MLP_3D.forward op for op: under torch.autocast on the CPU it reproduces
    the reference's own 16-bit outputs and gradients (tests/golden/mlp_autocast.npz, made by the unmodified reference),
    so it can stand in for "the reference at 16 bit" at sizes no fixture holds."""
    g = load_golden("mlp_autocast.npz")
    p = {k: v.clone().requires_grad_() for k, v in mlp_params(g).items()}
    S = int(g["S"])
    f = g["feat"].clone().requires_grad_()
    drep = g["dirs"][:, None, :].repeat(1, S, 1).reshape(-1, g["dirs"].shape[-1])
    with torch.autocast("cpu", dtype=fmt or torch.bfloat16, enabled=fmt is not None):
        out = port.mlp_forward(p, f, drep)
    out.float().backward(g["dout"])
    tol = dict(rtol=1e-6, atol=1e-7) if fmt is None else dict(rtol=0, atol=0)
    assert torch.allclose(out.detach().float(), g[f"{tag}__out"], **tol)
    assert torch.allclose(f.grad, g[f"{tag}__dfeat"], **tol)
    for k, v in p.items():
        assert torch.allclose(v.grad, g[f"{tag}__grad__" + k.replace(".", "__")], **tol), k
