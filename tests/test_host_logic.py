"""CPU: host-side logic of the package that needs no device -- run grouping of the fused optimiser, the packed batch
layout of the graphed step, the pixel order of the ray ids."""
import torch

from human_body_reconstruction_b200 import optim
from human_body_reconstruction_b200.graph import GraphedStep, default_loss


def _views(flat, shapes):
    out, o = [], 0
    for s in shapes:
        n = int(torch.tensor(s).prod())
        out.append(flat[o:o + n].view(s))
        o += n
    return out


def test_runs_merge_adjacent_parameters_only_when_grads_are_adjacent_too():
    shapes = [(4, 2), (3,), (5, 2), (2,)]
    n = sum(int(torch.tensor(s).prod()) for s in shapes)
    flat, gflat = torch.zeros(n), torch.zeros(n)
    ps = [torch.nn.Parameter(v) for v in _views(flat, shapes)]
    for p, g in zip(ps, _views(gflat, shapes)):
        p.grad = g
    assert optim._runs(ps) == [(0, 4, n)]                                   # one launch for the whole buffer
    ps[2].grad = torch.zeros(5, 2)                                          # a gradient living elsewhere splits the run
    assert optim._runs(ps) == [(0, 2, 11), (2, 1, 10), (3, 1, 2)]
    lone = torch.nn.Parameter(torch.zeros(7))
    lone.grad = torch.zeros(7)
    assert optim._runs([lone] + ps[:2]) == [(0, 1, 7), (1, 2, 11)]
    assert optim._runs([]) == []


def test_pack_batch_layout_matches_static_buffer_views():
    R = 5
    o, d = torch.arange(3 * R).float().view(R, 3), 100 + torch.arange(3 * R).float().view(R, 3)
    n, gt = 200 + torch.arange(R).float().view(R, 1), 300 + torch.arange(3 * R).float().view(R, 3)
    flat = GraphedStep.pack_batch(o, d, n, gt)
    assert flat.shape == (10 * R,)
    assert torch.equal(flat[0:3 * R].view(R, 3), o) and torch.equal(flat[3 * R:6 * R].view(R, 3), d)
    assert torch.equal(flat[6 * R:7 * R].view(R, 1), n) and torch.equal(flat[7 * R:].view(R, 3), gt)


def test_default_loss_is_mse_plus_mse():
    g = torch.Generator().manual_seed(0)
    Cr, Cf, gt = torch.rand(6, 3, generator=g), torch.rand(6, 3, generator=g), torch.rand(6, 3, generator=g)
    mse = torch.nn.functional.mse_loss
    assert torch.equal(default_loss(Cr, Cf, gt), mse(Cr, gt) + mse(Cf, gt))      # train_hash2.py:221
    assert torch.equal(default_loss(Cr, Cr, gt), mse(Cr, gt) + mse(Cr, gt))      # x + x == 2 x exactly


def test_same_seed_construction_reproduces_the_reference_parameters():
    """Drop-in contract: same constructor arguments + same torch seed -> the same initial parameters as the
    reference, i.e. the modules consume the RNG in the reference's order (hash_encoding.py:30-32: one nn.Embedding +
    uniform_(-1e-4, 1e-4) per level; test_hash.py:27-50: sigma net then colour net).  The fixture grid.npz was built by
    the reference with torch.manual_seed(14); HashEncoder(...), MLP_3D(...) in that order (oracle/make_golden.py: build)."""
    import human_body_reconstruction_b200 as h
    from conftest import load_golden, mlp_params
    g = load_golden("grid.npz")
    L, T, F = g["tables"].shape
    torch.manual_seed(14)
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=L, F=F, T=T, dim=3, mu=g["mu"], sigma=g["sigma"], device="cpu")
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=L, F=F, d_view=24, max_bound=g["max_bound"], min_bound=g["min_bound"])
    tables = torch.stack([e.weight.detach() for e in enc.Embedding_list])
    assert torch.equal(tables * 5e3, g["tables"])                             # the fixture scaled its tables by 5e3
    assert [float(s) for s in g["scales"]] == enc.level_scales()
    want = mlp_params(g, "mlp__")
    sd = {k: v for k, v in mlp.state_dict().items()}
    assert set(sd) == set(want)
    for k in want:
        assert torch.equal(sd[k].cpu(), want[k]), k
    assert list(enc.state_dict().keys()) == [f"Embedding_list.{i}.weight" for i in range(L)]


def test_sdf_compositor_accepts_column_views_of_the_packed_field_output_without_copies():
    """ops._sample_stride: the SDF compositor addresses rgb / sdf as base + (r*S+s)*stride, so the (R,S,3) / (R,S) column
    views of the MLP's packed (R*S,4) output (what vol_renderer.py:213-216 hands calc_color) go in as they are (stride 4),
    contiguous tensors with strides 3 / 1, anything else is copied."""
    from human_body_reconstruction_b200.ops import _sample_stride, _sdf_views
    R, S = 5, 7
    mo = torch.randn(R * S, 4)
    rgb, sdf = mo[..., 0:3].reshape(R, S, -1), mo[..., 3:4].reshape(R, S)
    assert _sample_stride(rgb, 3) == 4 and _sample_stride(sdf, 1) == 4
    r2, rs, s2, ss = _sdf_views(rgb, sdf)
    assert r2.data_ptr() == mo.data_ptr() and s2.data_ptr() == mo.data_ptr() + 12 and (rs, ss) == (4, 4)
    assert _sample_stride(torch.randn(R, S, 3), 3) == 3 and _sample_stride(torch.randn(R, S), 1) == 1
    assert _sample_stride(torch.randn(S, R, 3).transpose(0, 1), 3) is None and _sample_stride(torch.randn(S, R).t(), 1) is None
    r3, rs3, s3, ss3 = _sdf_views(torch.randn(S, R, 3).transpose(0, 1).half(), torch.randn(S, R).t())
    assert r3.is_contiguous() and r3.dtype == torch.float32 and s3.is_contiguous() and (rs3, ss3) == (3, 1)
    for shape in ((1, 1), (1, S), (R, 1)):                                 # degenerate shapes keep a usable stride
        assert _sample_stride(torch.randn(*shape, 3), 3) == 3 and _sample_stride(torch.randn(*shape), 1) == 1


def test_mlp_bounds_for_the_eikonal_stencil_follow_the_tensors():
    """MLP_3D._bounds_host: min_bound / max_bound (test_hash.py:25-26) as three host floats each for the stencil kernel's
    clamp; scalars broadcast; re-read when the tensors change."""
    import human_body_reconstruction_b200 as h
    m = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=torch.tensor([1., 2., 3.]), min_bound=torch.tensor([-1., -2., -3.]))
    assert m._bounds_host() == ([-1.0, -2.0, -3.0], [1.0, 2.0, 3.0])
    m.max_bound[1] = 5.0
    assert m._bounds_host()[1] == [1.0, 5.0, 3.0]
    assert h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24)._bounds_host() == ([-1.0] * 3, [1.0] * 3)


def test_sdf_route_selection_is_by_module_type():
    """Volume_Renderer._native_sdf: the dedicated SDF route needs the native encoder / direction encoder / MLP_3D AND the
    reference's own VarModel (whose parameter the compositing kernel reads); anything else keeps the reference's data
    flow (_generic), where calc_color decides per call."""
    import human_body_reconstruction_b200 as h
    enc = h.HashEncoder(N_min=16, N_max=64.0, L=4, F=2, T=64, dim=3, mu=torch.zeros(3), sigma=torch.tensor(1.0), device="cpu")
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=4, F=2, d_view=24)
    pe = h.PositionalEncoder(3, 4)

    def renderer(var, pos=enc):
        return h.Volume_Renderer(H=4, W=4, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device="cpu", Pos_encode=pos,
                                 Dir_encode=pe, max_dim=16, sigma_val=torch.tensor(1.0), mu=torch.zeros(3), use_sdf=True, var_model=var)

    class OtherVar(h.helper.VarModel):
        def forward(self, x):
            return torch.sigmoid(2 * x * self.b)

    vr = renderer(h.helper.VarModel())
    assert vr._native_sdf(mlp) is mlp and vr._native_sdf(torch.nn.DataParallel(mlp, device_ids=[0]) if torch.cuda.is_available() else mlp) is mlp
    assert vr._native(mlp) is None                                     # never the NeRF-mode route
    assert renderer(OtherVar())._native_sdf(mlp) is None               # a different phi: tensor expressions
    assert renderer(h.helper.VarModel(), pos=torch.nn.Identity())._native_sdf(mlp) is None
    assert renderer(h.helper.VarModel())._native_sdf(torch.nn.Linear(4, 4)) is None
    assert vr.sdf_native is True and h.helper.SDF_KERNELS is True
