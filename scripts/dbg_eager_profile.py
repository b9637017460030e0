"""Diagnostic (GPU box): cProfile of the eager (non-graph) bench step -- where does the host time go?"""
import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import human_body_reconstruction_b200 as hbr
dev = torch.device("cuda", 0)
H = W = 800
c2w, K = bench.make_cameras(100, 0), bench.intrinsics(H, W)
mx, mn = bench.scene_bbox(c2w, K, H, W, 2.0, 6.0)
sigma = ((mx - mn) ** 2).sum().sqrt()
torch.manual_seed(0)
enc = hbr.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=2**19, dim=3, mu=mn.to(dev), sigma=sigma.to(dev))
mlp = hbr.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=mx, min_bound=mn)
enc, mlp = enc.to(dev), mlp.to(dev)
nerf = torch.nn.DataParallel(mlp, device_ids=[0])
vr = hbr.Volume_Renderer(H=H, W=W, K=K, near=torch.tensor(2.0), far=torch.tensor(6.0), device=dev, Pos_encode=enc,
                         Dir_encode=hbr.PositionalEncoder(3, 4), max_dim=1024, sigma_val=sigma, mu=mn)
batches = [tuple(t.to(dev) for t in b) for b in bench.make_batches(c2w, K, H, W, 4096, 4, 100)]
params = list(enc.parameters()) + list(mlp.parameters())
def step(b):
    o, d, n, gt = b
    for p in params: p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        Cr, Cf, _ = vr.vol_render(nerf, d, o, num_samples=128, update_mask=False, dir_norm=n, hierarchical=False)
        loss = torch.nn.functional.mse_loss(Cr, gt) + torch.nn.functional.mse_loss(Cf, gt)
    loss.backward()
for k in range(10): step(batches[k % 4])
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for k in range(100): step(batches[k % 4])
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"eager: host {1e3*(t1-t0)/100:.3f} ms/step, incl. drain {1e3*(t2-t0)/100:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for k in range(100): step(batches[k % 4])
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
