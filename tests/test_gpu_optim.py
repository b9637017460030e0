"""GPU tests (-m gpu) of the fused Adam / AdamW step against torch.optim.Adam / AdamW -- the optimisers
train_hash2.py:141-142 constructs -- evaluated on the CPU on identical parameters and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("adamw,wd", [(False, 0.0), (True, 0.01), (False, 0.02)])
def test_fused_adam_matches_torch(adamw, wd):
    from human_body_reconstruction_b200.optim import FusedAdam
    torch.manual_seed(3)
    shapes = [(1024, 2)] * 5 + [(64, 32), (64,), (7,)]
    flat = torch.randn(sum(torch.Size(s).numel() for s in shapes))
    # reference parameters (CPU) and ours: views of ONE flat CUDA buffer (as HashEncoder / MLP_3D lay them out)
    ref, ours, o = [], [], 0
    flat_dev = flat.to(DEV)
    for s in shapes:
        k = torch.Size(s).numel()
        ref.append(torch.nn.Parameter(flat[o:o + k].view(s).clone()))
        ours.append(torch.nn.Parameter(flat_dev[o:o + k].view(s)))
        o += k
    kw = dict(lr=0.05, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    opt_ref = (torch.optim.AdamW if adamw else torch.optim.Adam)(ref, **kw)
    opt = FusedAdam(ours, decoupled=adamw, **kw)
    from human_body_reconstruction_b200 import _lib
    for it in range(5):
        g = torch.randn_like(flat) * (0.1 + it)
        gflat = g.to(DEV)
        o = 0
        for pr, po, s in zip(ref, ours, shapes):
            k = torch.Size(s).numel()
            pr.grad = g[o:o + k].view(s).clone()
            po.grad = gflat[o:o + k].view(s)          # adjacent views of one flat gradient buffer
            o += k
        _lib.STATS.reset()
        opt_ref.step()
        opt.step()
        assert _lib.STATS.calls.get("hbr_adam_step") == 1, "adjacent parameters must be merged into one launch"
        for pr, po in zip(ref, ours):
            assert rel(po, pr) < 2e-6
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) >= {"exp_avg", "exp_avg_sq", "step"} and sd["state"][0]["step"] == 5
    assert rel(opt.state[ours[0]]["exp_avg"], opt_ref.state[ref[0]]["exp_avg"]) < 2e-6
    assert rel(opt.state[ours[0]]["exp_avg_sq"], opt_ref.state[ref[0]]["exp_avg_sq"]) < 2e-6


def test_fused_adam_grad_scaler_protocol_and_separate_buffers():
    from human_body_reconstruction_b200.optim import FusedAdam
    torch.manual_seed(1)
    a = torch.nn.Parameter(torch.randn(1000, device=DEV))
    b = torch.nn.Parameter(torch.randn(33, device=DEV))          # separate allocations: two launches
    ra, rb = (torch.nn.Parameter(t.detach().cpu().clone()) for t in (a, b))
    opt, ref = FusedAdam([a, b], lr=0.01), torch.optim.Adam([ra, rb], lr=0.01)
    ga, gb = torch.randn(1000), torch.randn(33)
    scale = 1024.0
    a.grad, b.grad = (ga * scale).to(DEV), (gb * scale).to(DEV)
    ra.grad, rb.grad = ga.clone(), gb.clone()
    before = a.detach().clone()
    opt.step(inv_scale=1.0 / scale, found_inf=torch.ones(1, device=DEV))       # inf found: the step is skipped
    assert torch.equal(a.detach(), before)
    opt2 = FusedAdam([a, b], lr=0.01)
    opt2.step(inv_scale=1.0 / scale, found_inf=torch.zeros(1, device=DEV))
    ref.step()
    assert rel(a, ra) < 2e-6 and rel(b, rb) < 2e-6


def test_fused_adam_on_encoder_tables_single_launch():
    """The encoder's 16 level tables are adjacent slices of one buffer and so are the gradients its backward returns:
    the whole 16 x T x F table is updated by one kernel."""
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200 import _lib
    torch.manual_seed(0)
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=2 ** 12, dim=3, mu=torch.zeros(3, device=DEV),
                        sigma=torch.tensor(1.0, device=DEV)).to(DEV)
    w0 = torch.stack([e.weight.detach().clone() for e in enc.Embedding_list]).cpu()
    x = torch.rand(4096, 3, device=DEV)
    enc(x).square().sum().backward()
    g = torch.stack([e.weight.grad for e in enc.Embedding_list]).cpu()
    ref_p = torch.nn.Parameter(w0.clone())
    ref_p.grad = g.clone()
    torch.optim.Adam([ref_p], lr=0.05).step()
    opt = h.optim.FusedAdam(enc.Embedding_list.parameters(), lr=0.05)
    _lib.STATS.reset()
    opt.step()
    assert _lib.STATS.calls.get("hbr_adam_step") == 1
    got = torch.stack([e.weight.detach() for e in enc.Embedding_list]).cpu()
    assert rel(got, ref_p) < 2e-6
    # the encoder still sees one flat table (the update was in place)
    y = enc(x)
    assert torch.isfinite(y).all()
