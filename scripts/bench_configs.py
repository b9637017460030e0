"""Secondary measurements (GPU box) at BASELINE.json configs[3] and configs[4] -- not bench.py lines, recorded in profiles/."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import human_body_reconstruction_b200 as h
dev = "cuda"
MU, MAXB = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
SIGMA = ((MAXB - MU) ** 2).sum().sqrt()

def build(T, scale):
    torch.manual_seed(0)
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=MU.to(dev), sigma=SIGMA.to(dev))
    with torch.no_grad():
        for e in enc.Embedding_list: e.weight.mul_(scale)
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=MAXB, min_bound=MU)
    return enc.to(dev), mlp.to(dev)

def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), r

out = {}
# ---- configs[3]: 512^3 density grid + marching cubes ----
enc, mlp = build(2 ** 19, 2e5)
mn, mx = MU.double().tolist(), MAXB.double().tolist()
res = 512
ms, dens = ev_time(lambda: h.mesh.density_grid(enc, mlp, None, mn, mx, res))
iso = float(dens.median())
ms_c, cnt = ev_time(lambda: h.mesh.marching_cubes_counts(dens, iso))
ms_e, vf = ev_time(lambda: h.mesh.marching_cubes(dens, iso), reps=2)
out["grid_512"] = {"points": res ** 3, "density_ms": ms, "Mpts_per_s": res ** 3 / ms / 1e3, "algorithmic_GBps": 1034 * res ** 3 / ms / 1e6,
                   "mc_count_ms": ms_c, "mc_emit_ms": ms_e, "vertices": cnt[0], "triangles": cnt[1], "fp32 MLP": True}
pe = h.PositionalEncoder(3, 4)
ms4, _ = ev_time(lambda: h.mesh.density_grid(enc, mlp, pe, mn, mx, 256), reps=2)
out["grid_256_rgb_density"] = {"points": 256 ** 3, "ms": ms4, "Mpts_per_s": 256 ** 3 / ms4 / 1e3,
                               "reference_cpu_s": 185.0, "note": "reference: nerf2mesh.py 256^3 on 8 CPU threads (SURVEY section 6 probe, T=2^14)"}
del dens, vf
torch.cuda.empty_cache()
# ---- configs[4]: T = 2^22, 256 samples/ray, hierarchical ----
enc, mlp = build(2 ** 22, 1e4)
vr = h.Volume_Renderer(H=1080, W=1920, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device=dev,
                       Pos_encode=enc, Dir_encode=pe, max_dim=1024, sigma_val=SIGMA, mu=MU)
R = 4096
g = torch.Generator().manual_seed(1)
ro = (torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.05 * torch.randn(R, 3, generator=g)).to(dev)
rd = torch.nn.functional.normalize(-ro.cpu() + 0.5 * torch.randn(R, 3, generator=g), dim=-1).to(dev)
dn = (1 + 0.2 * torch.rand(R, 1, generator=g)).to(dev)
gt = torch.rand(R, 3, generator=g).to(dev)
params = list(enc.parameters()) + list(mlp.parameters())
gs = h.graph.GraphedStep(vr, mlp, params, R, 256, True, dev).capture()
gs.load(ro, rd, dn, gt)
ms5, _ = ev_time(lambda: gs(), reps=5)
pts = R * 768
out["human_config_step"] = {"rays": R, "points_per_ray": 768, "T": 2 ** 22, "ms_per_step": ms5, "rays_per_s": R / ms5 * 1e3,
                            "step_algorithmic_GBps": 2792 * pts / ms5 / 1e6, "launch": "cuda graph replay, L2 warm"}
# ---- 8f row 1: optimiser step over the 16 x 2^19 x 2 table + the MLP (train_hash2.py:141-142,227-228) ----
del gs, enc, mlp, vr
torch.cuda.empty_cache()
enc, mlp = build(2 ** 19, 1e4)
for p in list(enc.parameters()) + list(mlp.parameters()):
    p.grad = torch.randn_like(p) * 1e-3
# gradients as this package's backward delivers them: views of one flat buffer
gflat = torch.randn(16, 2 ** 19, 2, device=dev) * 1e-3
for i, e in enumerate(enc.Embedding_list):
    e.weight.grad = gflat[i]
def timed_opt(make):
    o1, o2 = make()
    def stepf():
        o1.step(); o2.step()
    return ev_time(stepf, reps=5)[0]
n_par = sum(p.numel() for p in enc.parameters()) + sum(p.numel() for p in mlp.parameters())
res = {}
res["torch_adam_default_ms"] = timed_opt(lambda: (torch.optim.Adam(enc.Embedding_list.parameters(), lr=.05), torch.optim.AdamW(mlp.parameters(), lr=.005)))
res["torch_adam_fused_ms"] = timed_opt(lambda: (torch.optim.Adam(enc.Embedding_list.parameters(), lr=.05, fused=True), torch.optim.AdamW(mlp.parameters(), lr=.005, fused=True)))
res["hbr_fused_adam_ms"] = timed_opt(lambda: (h.optim.FusedAdam(enc.Embedding_list.parameters(), lr=.05), h.optim.FusedAdamW(mlp.parameters(), lr=.005)))
res["params"] = n_par
res["hbr_GBps"] = 28 * n_par / res["hbr_fused_adam_ms"] / 1e6
res["roofline_frac"] = res["hbr_GBps"] / 6544.7
out["optimizer_step"] = res
print(json.dumps(out))
