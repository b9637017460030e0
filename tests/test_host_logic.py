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
