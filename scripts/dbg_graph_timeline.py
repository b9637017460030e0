"""Diagnostic (GPU box): device timeline of one CUDA-graph replay of the bench step (kernel durations and gaps)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import human_body_reconstruction_b200 as hbr
from human_body_reconstruction_b200.graph import GraphedStep
dev = torch.device("cuda", 0)
H = W = 800
c2w, K = bench.make_cameras(100, 0), bench.intrinsics(H, W)
mx, mn = bench.scene_bbox(c2w, K, H, W, 2.0, 6.0)
sigma = ((mx - mn) ** 2).sum().sqrt()
torch.manual_seed(0)
enc = hbr.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=2**19, dim=3, mu=mn.to(dev), sigma=sigma.to(dev))
with torch.no_grad():
    for e in enc.Embedding_list: e.weight.mul_(1e4)
mlp = hbr.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=mx, min_bound=mn)
enc, mlp = enc.to(dev), mlp.to(dev)
nerf = torch.nn.DataParallel(mlp, device_ids=[0])
vr = hbr.Volume_Renderer(H=H, W=W, K=K, near=torch.tensor(2.0), far=torch.tensor(6.0), device=dev, Pos_encode=enc,
                         Dir_encode=hbr.PositionalEncoder(3, 4), max_dim=1024, sigma_val=sigma, mu=mn)
batches = [tuple(t.to(dev) for t in b) for b in bench.make_batches(c2w, K, H, W, 4096, 4, 100)]
params = list(enc.parameters()) + list(mlp.parameters())
gs = GraphedStep(vr, nerf, params, 4096, 128, False, dev).capture()
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
for k in range(5): gs(*batches[k % 4])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for k in range(3):
        flush.zero_()
        gs(*batches[k % 4])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# last replay = events after the last big fill (flush)
idx = [i for i, e in enumerate(ev) if "FillFunctor<float>" in e.name and (e.time_range.end - e.time_range.start) > 50]
seg = ev[idx[-1] + 1:]
t0 = seg[0].time_range.start; last = None; tot = 0; gaps = 0
for e in seg:
    d = e.time_range.end - e.time_range.start
    gap = (e.time_range.start - last) if last is not None else 0
    print(f"{e.name[:64]:64s} start {e.time_range.start - t0:8.1f} dur {d:8.1f} gap {gap:7.1f}")
    last = max(last or 0, e.time_range.end); tot += d; gaps += max(gap, 0)
print(f"span {last - t0:.1f} us, sum of kernel durations {tot:.1f} us, sum of positive gaps {gaps:.1f} us, kernels {len(seg)}")
