"""Diagnostic (GPU box): where does a bench step's time go?  host enqueue time vs device time, kernel gaps."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import human_body_reconstruction_b200 as hbr

class A: pass
args = A(); args.res=800; args.views=100; args.near=2.0; args.far=6.0; args.hash_size=19; args.max_res=2048.0
args.samples=128; args.rays=4096; args.hierarchical=False
dev = torch.device("cuda", 0)
H = W = args.res
c2w, K = bench.make_cameras(args.views, 0), bench.intrinsics(H, W)
mx, mn = bench.scene_bbox(c2w, K, H, W, args.near, args.far)
sigma = ((mx - mn) ** 2).sum().sqrt()
torch.manual_seed(0)
enc = hbr.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=2**19, dim=3, mu=mn.to(dev), sigma=sigma.to(dev))
with torch.no_grad():
    for e in enc.Embedding_list: e.weight.mul_(1e4)
mlp = hbr.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=mx, min_bound=mn)
enc, mlp = enc.to(dev), mlp.to(dev)
nerf = torch.nn.DataParallel(mlp, device_ids=[0])
pe = hbr.PositionalEncoder(3, 4)
vr = hbr.Volume_Renderer(H=H, W=W, K=K, near=torch.tensor(2.0), far=torch.tensor(6.0), device=dev, Pos_encode=enc, Dir_encode=pe,
                         max_dim=1024, sigma_val=sigma, mu=mn)
batches = [tuple(t.to(dev) for t in b) for b in bench.make_batches(c2w, K, H, W, args.rays, 4, 100)]
params = list(enc.parameters()) + list(mlp.parameters())
def step(b):
    o, d, n, gt = b
    for p in params: p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        Cr, Cf, _ = vr.vol_render(nerf, d, o, num_samples=128, update_mask=False, dir_norm=n, hierarchical=False)
        loss = torch.nn.functional.mse_loss(Cr, gt) + torch.nn.functional.mse_loss(Cf, gt)
    loss.backward()
    return loss
for k in range(5): step(batches[k % 4])
torch.cuda.synchronize()
# host enqueue time vs device time
N = 20
t0 = time.perf_counter()
for k in range(N): step(batches[k % 4])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/N:.3f} ms/step; total incl. drain {1e3*(t2-t0)/N:.3f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for k in range(3): step(batches[k % 4])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
last = None
print("---- device timeline of the last step (name, start us rel, dur us, gap before us)")
t_first = ev[0].time_range.start
n = len(ev) // 3
for e in ev[-n:]:
    gap = (e.time_range.start - last) if last is not None else 0
    print(f"{e.name[:60]:60s} {e.time_range.start - t_first:10.1f} {e.time_range.end - e.time_range.start:9.1f} {gap:9.1f}")
    last = e.time_range.end
