"""BASELINE configs[3] in isolation (the command the grid ncu capture profiles): density grid at --res over the scene bounds
+ marching-cubes count.  Prints ms per pass (CUDA events)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import human_body_reconstruction_b200 as hbr

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cuda-core-mlp", action="store_true")
a = ap.parse_args()
dev = "cuda"
mn, mx = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
sigma = ((mx - mn) ** 2).sum().sqrt()
torch.manual_seed(1)
enc = hbr.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=2 ** 19, dim=3, mu=mn.to(dev), sigma=sigma.to(dev))
with torch.no_grad():
    for l, e in enumerate(enc.Embedding_list):
        e.weight.mul_(2e5 if l < 4 else 2e3)
mlp = hbr.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=mx, min_bound=mn)
enc, mlp = enc.to(dev), mlp.to(dev)
lo, hi = mn.double().tolist(), mx.double().tolist()
for r in range(a.reps):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    dens = hbr.mesh.density_grid(enc, mlp, None, lo, hi, a.res, chunk=1 << 19, cuda_core_mlp=a.cuda_core_mlp)
    e1.record()
    c = hbr.ops.mc_count(dens, float(dens[0, 0, 0]) if r == 0 else iso)
    e2.record()
    torch.cuda.synchronize()
    iso = float(dens.float().median())
    print(f"res {a.res}: density {e0.elapsed_time(e1):.2f} ms, mc_count {e1.elapsed_time(e2):.3f} ms, counts {c.tolist()}")
