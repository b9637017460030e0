"""on-device ray generation (hbr_ray_gen / hbr_ray_bbox, DeviceRayDataset) against the fixtures
generated from the reference's get_od / find_bounding_box and against the oracle on random rigs.
Tolerance: 2e-6 absolute on unit directions / 1e-6 relative elsewhere -- the 3-term rotation d = R @ dirs is a BLAS call
in the reference whose summation order is unspecified; everything else (pixel -> camera coordinates, the division by
the norm) follows the reference's separate roundings."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(a, b, atol=2e-6, rtol=1e-6):
    return torch.allclose(a.cpu(), b.cpu(), atol=atol, rtol=rtol)


def test_ray_gen_matches_reference_fixture():
    import human_body_reconstruction_b200 as h
    g = load_golden("rays.npz")
    H, W, K, c2w = int(g["H"]), int(g["W"]), g["K"], g["c2w"]
    o, d, n, gt = h.ops.ray_gen(c2w.to(DEV), H, W, K)
    V = c2w.shape[0]
    assert gt is None and o.shape == (V * H * W, 3)
    assert torch.equal(o.cpu().view(V, H * W, 3), g["rays_o"])                 # origins are copies: exact
    assert _close(d.view(V, H * W, 3), g["rays_d"]) and _close(n.view(V, H * W, 1), g["dir_norm"])
    # the drop-in helper routes CUDA poses through the same kernel and keeps the reference's shapes
    o2, d2, n2 = h.helper.get_od(H, W, K, c2w.to(DEV))
    assert o2.shape == g["rays_o"].shape and d2.shape == g["rays_d"].shape and n2.shape == g["dir_norm"].shape
    assert torch.equal(d2.reshape(-1, 3), d)
    # bounding box: the reference derives W = 2*K[0,2], H = 2*K[1,2] (helper.py:114-115) -> 6 x 4 here
    loader = [(None, c2w[i:i + 2], None) for i in (0, 2)]
    mx, mn = h.helper.find_bounding_box(loader, near=torch.tensor(2.0), far=torch.tensor(6.0), K=K, device=DEV)
    assert _close(mx, g["max_bound"], atol=1e-5) and _close(mn, g["min_bound"], atol=1e-5)


def _rig(V, seed):
    gen = torch.Generator().manual_seed(seed)
    c2w = torch.eye(4).repeat(V, 1, 1)
    q, _ = torch.linalg.qr(torch.randn(V, 3, 3, generator=gen))
    c2w[:, :3, :3] = q
    c2w[:, :3, 3] = torch.randn(V, 3, generator=gen) * 3
    return c2w, gen


@pytest.mark.parametrize("pix", ["u8", "f32"])
def test_ray_gen_random_ids_and_pixels_vs_oracle(pix):
    import human_body_reconstruction_b200 as h
    V, H, W = 5, 31, 45
    c2w, gen = _rig(V, 3)
    K = torch.tensor([[50.3, 0, 22.5], [0, 49.1, 15.5], [0, 0, 1]])
    if pix == "u8":
        img = torch.randint(0, 256, (V, H, W, 3), generator=gen, dtype=torch.uint8)
        want_gt = img.float() / 255                                           # torchvision ToTensor
    else:
        img = torch.rand(V, H, W, 3, generator=gen)
        want_gt = img
    ids = torch.randint(0, V * H * W, (4097,), generator=gen)
    ids[:3] = torch.tensor([0, V * H * W - 1, W])                             # first, last, start of row 1
    o, d, n = port.get_od(H, W, K, c2w)
    o, d, n = o.reshape(-1, 3), d.reshape(-1, 3), n.reshape(-1, 1)
    go, gd, gn, ggt = h.ops.ray_gen(c2w.to(DEV), H, W, K, ray_ids=ids.to(DEV), images=img.to(DEV), check_ids=True)
    assert torch.equal(go.cpu(), o[ids]) and _close(gd, d[ids]) and _close(gn, n[ids])
    assert torch.equal(ggt.cpu(), want_gt.reshape(-1, 3)[ids])
    assert float((gd.norm(dim=-1) - 1).abs().max()) < 1e-6
    with pytest.raises(IndexError):
        h.ops.ray_gen(c2w.to(DEV), H, W, K, ray_ids=torch.tensor([V * H * W], device=DEV), check_ids=True)


def test_device_ray_dataset_epoch_and_bbox():
    import human_body_reconstruction_b200 as h
    V, H, W = 3, 20, 24
    c2w, gen = _rig(V, 5)
    K = torch.from_numpy(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]]))
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = 27, 27, W // 2, H // 2
    ids_img = torch.arange(V * H * W, dtype=torch.float32).view(V, H, W, 1).repeat(1, 1, 1, 3)
    ds = h.DeviceRayDataset(ids_img.permute(0, 3, 1, 2), c2w, K, device=DEV, batch_size=500,
                            generator=torch.Generator(device=DEV).manual_seed(1))
    assert ds.n_rays == V * H * W and len(ds) == 3
    o_all, d_all, n_all = port.get_od(H, W, K, c2w)
    seen = []
    for k, (o, d, n, gt) in enumerate(ds):
        assert o.shape[0] == (500 if k < 2 else V * H * W - 1000)              # last batch partial, like DataLoader
        ids = gt[:, 0].long().cpu()
        assert _close(d, d_all.reshape(-1, 3)[ids]) and torch.equal(o.cpu(), o_all.reshape(-1, 3)[ids])
        seen.append(ids)
    seen = torch.cat(seen)
    assert not torch.equal(seen, torch.arange(V * H * W))                      # shuffled ...
    assert torch.equal(seen.sort().values, torch.arange(V * H * W))           # ... and every ray exactly once
    o, d, n, gt = ds.sample(64)
    assert o.shape == (64, 3) and gt.shape == (64, 3)
    mx, mn = ds.bounding_box(2.0, 6.0)
    wmx, wmn = port.bounding_box(c2w, K, 2.0, 6.0)
    assert _close(mx, wmx, atol=1e-5) and _close(mn, wmn, atol=1e-5)


def test_graphed_step_draws_its_own_batches():
    """graph.GraphedStep(source=DeviceRayDataset): the sampler is captured with the step, every replay sees a fresh batch
    (the device generator advances inside the graph) and produces a finite loss and gradients."""
    import human_body_reconstruction_b200 as h
    from human_body_reconstruction_b200.graph import GraphedStep
    V, H, W = 4, 16, 16
    c2w, gen = _rig(V, 7)
    c2w[:, :3, 3] = torch.nn.functional.normalize(c2w[:, :3, 3], dim=-1) * 4
    K = torch.tensor([[20.0, 0, 8], [0, 20.0, 8], [0, 0, 1]])
    img = torch.rand(V, H, W, 3, generator=gen)
    ds = h.DeviceRayDataset(img, c2w, K, device=DEV, batch_size=256)
    mx, mn = ds.bounding_box(2.0, 6.0)
    sigma = ((mx - mn) ** 2).sum().sqrt()
    torch.manual_seed(0)
    enc = h.HashEncoder(N_min=16, N_max=512.0, L=16, F=2, T=2 ** 14, dim=3, mu=mn, sigma=sigma).to(DEV)
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=mx, min_bound=mn).to(DEV)
    vr = h.Volume_Renderer(H=H, W=W, K=K, near=torch.tensor(2.0), far=torch.tensor(6.0), device=DEV, Pos_encode=enc,
                           Dir_encode=h.PositionalEncoder(3, 4), max_dim=64, sigma_val=sigma, mu=mn)
    params = list(enc.parameters()) + list(mlp.parameters())
    gs = GraphedStep(vr, mlp, params, 256, 32, False, DEV, source=ds).capture()
    losses = [float(gs().item()) for _ in range(4)]
    assert all(np.isfinite(losses)) and len(set(losses)) > 1                   # a different batch every replay
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)


def test_scene_on_disk_to_device_batches():
    """formats.load_scene -> DeviceRayDataset: the batches equal the reference's data path (dataset.py reader ->
    get_od with the truncated int64 K -> flattened TensorDataset rows, train_hash2.py:74-96) for the committed scene."""
    import os
    import human_body_reconstruction_b200 as h
    from conftest import GOLDEN
    g = load_golden("formats.npz")
    images, c2w, K, meta = h.formats.load_scene(os.path.join(GOLDEN, "scene_blender", "transforms_train.json"))
    H, W = int(meta["H"]), int(meta["W"])
    ds = h.DeviceRayDataset(images, c2w, K, device=DEV, batch_size=16, shuffle=False)
    o_all, d_all, n_all = port.get_od(H, W, K, c2w)
    gts = g["blender__images"].permute(0, 2, 3, 1).reshape(-1, 3)                     # train_hash2.py:81-82
    rows = [b for b in ds]
    o = torch.cat([b[0] for b in rows]).cpu()
    d = torch.cat([b[1] for b in rows]).cpu()
    n = torch.cat([b[2] for b in rows]).cpu()
    gt = torch.cat([b[3] for b in rows]).cpu()
    assert torch.equal(o, o_all.reshape(-1, 3)) and torch.equal(gt, gts)
    assert _close(d, d_all.reshape(-1, 3)) and _close(n, n_all.reshape(-1, 1))
