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


def test_fused_adam_device_scalars_skip_and_load_state_dict():
    """capturable mode (step count / lr / gradient scale / found_inf on the device): equals torch.optim.Adam; a step skipped
    through found_inf leaves parameters, moments AND the step count alone (torch's fused contract); load_state_dict after a
    step is honoured (the cached flat moments are rebuilt)."""
    from human_body_reconstruction_b200.optim import FusedAdam
    torch.manual_seed(2)
    flat = torch.randn(4096 + 64, device=DEV)
    ours = [torch.nn.Parameter(flat[:4096].view(64, 64)), torch.nn.Parameter(flat[4096:])]
    ref = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ours]
    opt, opt_ref = FusedAdam(ours, lr=0.02, capturable=True), torch.optim.Adam(ref, lr=0.02)
    found = torch.zeros(1, device=DEV)
    scale = torch.full((1,), 8.0, device=DEV)
    for it in range(6):
        g = torch.randn(4096 + 64) * 0.3
        gd = g.to(DEV)
        ours[0].grad, ours[1].grad = (gd[:4096] * 8.0).view(64, 64), gd[4096:] * 8.0       # scaled gradients, as under GradScaler
        ref[0].grad, ref[1].grad = g[:4096].view(64, 64).clone(), g[4096:].clone()
        skip = it == 2
        found.fill_(1.0 if skip else 0.0)
        before = [p.detach().clone() for p in ours]
        opt.grad_scale, opt.found_inf = scale, found                 # what GradScaler.step() sets before calling step()
        opt.step()
        del opt.grad_scale, opt.found_inf
        if skip:
            assert all(torch.equal(a, b) for a, b in zip(before, ours))
        else:
            opt_ref.step()
        if it == 3:
            opt.set_lr(0.005)
            opt_ref.param_groups[0]["lr"] = 0.005
    assert int(opt.state[ours[0]]["step"].item()) == 5               # six calls, one skipped
    for po, pr in zip(ours, ref):
        assert rel(po, pr) < 5e-6
    # load_state_dict after steps: the next step must use the LOADED moments
    sd = opt.state_dict()
    for st in sd["state"].values():
        st["exp_avg"] = torch.zeros_like(st["exp_avg"])
        st["exp_avg_sq"] = torch.zeros_like(st["exp_avg_sq"])
    opt.load_state_dict(sd)
    ours[0].grad, ours[1].grad = torch.ones(64, 64, device=DEV), torch.ones(64, device=DEV)
    opt.step()
    assert rel(opt.state[ours[0]]["exp_avg"], torch.full((64, 64), 0.1)) < 1e-6      # (1 - beta1) * g from ZERO moments


def test_optimizer_inside_the_captured_step():
    """graph.GraphedStep(optimizers=[FusedAdam(capturable=True)...]): one replay = render + loss + backward + both optimiser
    steps (train_hash2.py:218-239).  Parameters after k replays == k eager iterations on the same batches."""
    import copy
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200.graph import GraphedStep, default_loss
    from human_body_reconstruction_b200.optim import FusedAdam, FusedAdamW

    def build():
        torch.manual_seed(0)
        mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
        sigma = ((maxb - mu) ** 2).sum().sqrt()
        enc = h.HashEncoder(N_min=16, N_max=512.0, L=16, F=2, T=2 ** 12, dim=3, mu=mu.to(DEV), sigma=sigma.to(DEV))
        with torch.no_grad():
            for e in enc.Embedding_list:
                e.weight.mul_(1e3)
        mlp = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=maxb, min_bound=mu)
        enc, mlp = enc.to(DEV), mlp.to(DEV)
        vr = h.Volume_Renderer(H=8, W=8, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device=DEV,
                               Pos_encode=enc, Dir_encode=h.PositionalEncoder(3, 4), max_dim=64, sigma_val=sigma, mu=mu)
        opts = [FusedAdam(list(enc.Embedding_list.parameters()), lr=0.01, capturable=True),
                FusedAdamW(mlp.parameters(), lr=0.005, capturable=True)]
        return enc, mlp, vr, opts

    R, S = 64, 16
    g = torch.Generator().manual_seed(3)
    ro = (torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.05 * torch.randn(R, 3, generator=g)).to(DEV)
    rd = torch.nn.functional.normalize(-ro.cpu() + 0.5 * torch.randn(R, 3, generator=g), dim=-1).to(DEV)
    dn, gt = torch.ones(R, 1, device=DEV), torch.rand(R, 3, generator=g).to(DEV)
    t = torch.linspace(2.0, 6.0, S, device=DEV)

    enc, mlp, vr, opts = build()
    params = list(enc.parameters()) + list(mlp.parameters())
    orig = vr.vol_render
    vr.vol_render = lambda *a, **k: orig(*a, **{**k, "t": t})        # fixed depths: no RNG between the two runs
    gs = GraphedStep(vr, mlp, params, R, S, False, DEV, warmup=2, optimizers=opts)
    gs.load(ro, rd, dn, gt)                                           # the warm-up iterations of capture() train on the loaded batch
    gs.capture()
    for _ in range(3):
        gs(ro, rd, dn, gt)
    torch.cuda.synchronize()
    got = [p.detach().clone() for p in params]

    enc2, mlp2, vr2, opts2 = build()
    params2 = list(enc2.parameters()) + list(mlp2.parameters())
    for _ in range(2 + 3):                                            # the 2 warm-up iterations of capture() are real steps
        for p in params2:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            Cr, Cf, _ = vr2.vol_render(mlp2, rd, ro, num_samples=S, t=t, dir_norm=dn, hierarchical=False)
            loss = default_loss(Cr, Cf, gt)
        loss.backward()
        for o in opts2:
            o.step()
    for a, b in zip(got, params2):
        assert rel(a, b) < 1e-4


def test_fused_adam_under_torch_grad_scaler():
    """scaler.scale(loss).backward(); scaler.step(opt) (train_hash2.py:226-228): GradScaler hands grad_scale / found_inf
    tensors to the optimiser (no host sync), the fused step unscales inside the kernel; == torch.optim.Adam under the same scaler."""
    from human_body_reconstruction_b200.optim import FusedAdam
    torch.manual_seed(6)
    w0 = torch.randn(512, device=DEV)
    x = torch.randn(64, 512, device=DEV)
    res = []
    for fused in (True, False):
        w = torch.nn.Parameter(w0.clone())
        opt = FusedAdam([w], lr=0.01) if fused else torch.optim.Adam([w], lr=0.01)
        scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
        for it in range(4):
            opt.zero_grad(set_to_none=True)
            loss = ((x @ w) ** 2).mean() * (float("inf") if it == 1 else 1.0)       # iteration 1 overflows: skipped by both
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
        res.append(w.detach().clone())
    assert torch.isfinite(res[0]).all() and rel(res[0], res[1]) < 5e-6
