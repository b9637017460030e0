"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): rays sharded over 2 ranks, gradients all-reduced
inside backward by dist.GradAllReduce (NCCL) / dist.PeerGradAllReduce (one kernel over NVLink peer memory) == single-GPU
gradients of the concatenated batch -- single pass and hierarchical (two field passes inside one backward),
whole-buffer and chunk-overlapped exchanges."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(dev):
    import human_body_reconstruction_b200 as h
    torch.manual_seed(0)
    mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
    sigma = ((maxb - mu) ** 2).sum().sqrt()
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=2 ** 14, dim=3, mu=mu.to(dev), sigma=sigma.to(dev))
    with torch.no_grad():
        for e in enc.Embedding_list:
            e.weight.mul_(5e3)
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=maxb, min_bound=mu)
    enc, mlp = enc.to(dev), mlp.to(dev)
    vr = h.Volume_Renderer(H=8, W=8, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device=dev,
                           Pos_encode=enc, Dir_encode=h.PositionalEncoder(3, 4), max_dim=64, sigma_val=sigma, mu=mu)
    return enc, mlp, vr


def _batch():
    g = torch.Generator().manual_seed(11)
    R = 256
    ro = torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.05 * torch.randn(R, 3, generator=g)
    rd = torch.nn.functional.normalize(-ro + 0.5 * torch.randn(R, 3, generator=g), dim=-1)
    return (ro, rd, 1 + 0.2 * torch.rand(R, 1, generator=g), torch.rand(R, 3, generator=g), torch.rand(R, 32, generator=g),
            torch.linspace(2.0, 6.0, 32), torch.rand(32, generator=g))


def _grads(enc, mlp, vr, ro, rd, dn, gt, u, t, u_cand, dev, hier=False, amp=False):
    """hier=True: coarse + fine render, i.e. TWO passes through the encoder and the MLP inside one backward.
    amp=True: the autocast training path (tensor-core MLP kernels, positions formed from the rays in the hash kernels)."""
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        Cr, Cf, _ = vr.vol_render(mlp, rd.to(dev), ro.to(dev), num_samples=32, t=t.to(dev), dir_norm=dn.to(dev),
                                  hierarchical=hier, _u=u.to(dev), _u_cand=u_cand.to(dev))
        loss = torch.nn.functional.mse_loss(Cr, gt.to(dev)) + torch.nn.functional.mse_loss(Cf, gt.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    return (torch.stack([e.weight.grad for e in enc.Embedding_list]).cpu(),
            torch.cat([p.grad.reshape(-1) for p in mlp.parameters()]).cpu())


def _worker(rank, world, port_no, tmp, mode="nccl"):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port_no))
    from human_body_reconstruction_b200 import dist as hdist
    hdist.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    enc, mlp, vr = _build(dev)
    red = hdist.GradAllReduce(enc, mlp) if mode == "nccl" else hdist.PeerGradAllReduce(enc, mlp, transport=mode[5:])
    ro, rd, dn, gt, u, t, u_cand = _batch()
    sl = hdist.shard_rays(ro.shape[0], rank, world)
    # peer mode: several steps through the same persistent region; hier: two field passes inside one backward
    # "4s" / "8s": the streamed exchange (one scatter-add launch finishing 4 / 8 level chunks in order + one exchange launch
    # beside it); amp: the autocast training path (_FieldRaysFn), whose MLP gradient travels in the same exchange launch
    cases = [(1, False, False), (4, False, False), (1, True, False), (4, True, False)]
    if mode != "nccl":
        cases += [("4s", False, False), ("8s", True, False), ("8s", False, True), ("4s", True, True), (1, False, True)]
    for chunks, hier, amp in cases:
        streamed = isinstance(chunks, str)
        nch = int(chunks[:-1]) if streamed else chunks
        enc._grad_chunks = nch
        if mode != "nccl":
            red._nchunks, red.overlap, red.streamed = nch, nch > 1, streamed
            if red.overlap and red._side is None:
                red._side, red.ctas = torch.cuda.Stream(device=dev, priority=-1), 32
        for p in list(enc.parameters()) + list(mlp.parameters()):
            p.grad = None
        # the streamed autocast cases cover both producers: "8s" the level-major scatter-add (fed by level-major d(features)
        # from the MLP backward), "4s" the tile-major kernel launched chunk-major
        from human_body_reconstruction_b200 import vol_renderer as vrm
        vrm.STREAM_LEVEL_MAJOR = chunks == "8s"
        gt_tab, gt_mlp = _grads(enc, mlp, vr, ro[sl], rd[sl], dn[sl], gt[sl], u[sl], t, u_cand, dev, hier, amp)
        torch.save((gt_tab, gt_mlp), os.path.join(tmp, f"g{rank}_{chunks}_{int(hier)}_{int(amp)}.pt"))
    if mode != "nccl":
        assert not red.region.timed_out()
        red.region.raise_if_failed()
    import torch.distributed as tdist
    tdist.barrier()
    tdist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_gradients_equal_single_gpu(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    dev = torch.device("cuda", 0)
    enc, mlp, vr = _build(dev)
    _check(tmp_path, enc, mlp, vr, dev, bit_identical=False)


def _check(tmp_path, enc, mlp, vr, dev, bit_identical):
    seen = 0
    for amp in (False, True):
        for hier in (False, True):
            for p in list(enc.parameters()) + list(mlp.parameters()):
                p.grad = None
            want_tab, want_mlp = _grads(enc, mlp, vr, *_batch(), dev, hier, amp)
            # fp32: 1e-5 (the path's bar); autocast: the same 16-bit arithmetic per point on both sides, but the tiles (and
            # with them the fp32 summation order of the weight gradients inside the tensor cores) differ: 1e-4
            tol = 1e-4 if amp else 1e-5
            for chunks in (1, 4, "4s", "8s"):
                files = [os.path.join(tmp_path, f"g{rank}_{chunks}_{int(hier)}_{int(amp)}.pt") for rank in range(2)]
                if not all(os.path.exists(f) for f in files):
                    continue
                seen += 1
                got = [torch.load(f) for f in files]
                for rank, (tab, gm) in enumerate(got):
                    assert float((tab - want_tab).norm() / want_tab.norm()) < tol, (rank, chunks, hier, amp)
                    assert float((gm - want_mlp).norm() / want_mlp.norm()) < tol, (rank, chunks, hier, amp)
                if bit_identical:
                    assert torch.equal(got[0][0], got[1][0]) and torch.equal(got[0][1], got[1][1]), (chunks, hier, amp)
    assert seen >= 4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["ipc", "symm"])
def test_peer_memory_allreduce_equals_single_gpu(tmp_path, transport):
    """dist.PeerGradAllReduce: gradients written into the peer-mapped region by the backward kernels, averaged by
    hbr_allreduce_peer (one kernel over NVLink: peer loads/stores over CUDA-IPC mappings, or switch-side reduction
    through torch symmetric memory's NVLS multicast mapping), == single-GPU gradients; both ranks bit-identical."""
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), "peer-" + transport), nprocs=2, join=True)
    dev = torch.device("cuda", 0)
    enc, mlp, vr = _build(dev)
    _check(tmp_path, enc, mlp, vr, dev, bit_identical=True)


# ---- SDF mode sharded over 2 ranks ---------------------------------------------------------------------
def _build_sdf(dev):
    import human_body_reconstruction_b200 as h
    enc, mlp, _ = _build(dev)
    var = h.helper.VarModel().to(dev)
    mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
    sigma = ((maxb - mu) ** 2).sum().sqrt()
    vr = h.Volume_Renderer(H=8, W=8, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device=dev, Pos_encode=enc,
                           Dir_encode=h.PositionalEncoder(3, 4), max_dim=64, sigma_val=sigma, mu=mu, use_sdf=True, var_model=var)
    return enc, mlp, var, vr


def _sdf_grads(enc, mlp, var, vr, ro, rd, gt, t, dev, amp):
    """One SDF training step as train_hash2.py:218-226 forms it: two field passes inside one backward (the render and the
    eikonal stencil), three parameter sets (tables, MLP, VarModel.b)."""
    import human_body_reconstruction_b200 as h
    for p in list(enc.parameters()) + list(mlp.parameters()) + list(var.parameters()):
        p.grad = None
    with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
        Cr, Cf, norm = vr.vol_render(mlp, rd.to(dev), ro.to(dev), num_samples=32, t=t.to(dev), update_mask=False, dir_norm=1.0,
                                     hierarchical=False)
        loss = (torch.nn.functional.mse_loss(Cr, gt.to(dev)) + torch.nn.functional.mse_loss(Cf, gt.to(dev))
                + 0.1 * h.helper.eikonal_loss(norm))
    loss.backward()
    torch.cuda.synchronize()
    return (torch.stack([e.weight.grad for e in enc.Embedding_list]).cpu(),
            torch.cat([p.grad.reshape(-1) for p in mlp.parameters()]).cpu(), var.b.grad.detach().cpu().clone())


def _sdf_worker(rank, world, port_no, tmp, kind):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port_no))
    from human_body_reconstruction_b200 import dist as hdist
    hdist.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    enc, mlp, var, vr = _build_sdf(dev)
    red = hdist.attach_grad_allreduce(enc, mlp, kind=kind)             # "auto": the production default at 2 ranks
    hdist.attach_small_params(var.parameters())
    ro, rd, dn, gt, u, t, u_cand = _batch()
    sl = hdist.shard_rays(ro.shape[0], rank, world)
    for amp in (False, True):
        for step in range(2):                                          # twice through the same persistent buffers
            got = _sdf_grads(enc, mlp, var, vr, ro[sl], rd[sl], gt[sl], t, dev, amp)
        torch.save(got, os.path.join(tmp, f"sdf{rank}_{int(amp)}.pt"))
    region = getattr(red, "region", None)
    if region is not None:
        region.raise_if_failed()
    import torch.distributed as tdist
    tdist.barrier()
    tdist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind", ["auto", "nccl"])
def test_sdf_mode_sharded_gradients_equal_single_gpu(tmp_path, kind):
    """SDF mode over 2 ranks: the table / MLP gradients of BOTH field passes (render + eikonal stencil: _FieldRaysFn or
    HashEncoder/_MlpFn, then HashEncoder/_DensityFn) go through one exchange per backward, VarModel.b through
    dist.attach_small_params; == single-GPU gradients of the whole batch, identical on both ranks.  Tolerance 2e-3: the
    stencil's +-1/(2 eps) coefficients land on the same table entries with opposite signs, which amplifies the fp32
    summation-order noise of the atomics (the single-GPU fixture test carries the same bound)."""
    mp.spawn(_sdf_worker, args=(2, _free_port(), str(tmp_path), kind), nprocs=2, join=True)
    dev = torch.device("cuda", 0)
    enc, mlp, var, vr = _build_sdf(dev)
    ro, rd, dn, gt, u, t, u_cand = _batch()
    for amp in (False, True):
        want = _sdf_grads(enc, mlp, var, vr, ro, rd, gt, t, dev, amp)
        got = [torch.load(os.path.join(tmp_path, f"sdf{rank}_{int(amp)}.pt")) for rank in range(2)]
        for rank in range(2):
            for name, a, b in zip(("tables", "mlp", "b"), got[rank], want):
                err = float((a - b).norm() / b.norm())
                assert err < 2e-3, (kind, amp, rank, name, err)
        for a, b in zip(got[0], got[1]):
            assert torch.equal(a, b), (kind, amp)
