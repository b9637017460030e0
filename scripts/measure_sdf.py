"""SDF mode on the GPU box: measured errors of the dedicated kernels on the reference fixture and the time
of one SDF training step (4096 rays x 128 samples, L=16 F=2 T=2^19) on the three routes -- native (Volume_Renderer's SDF
route on csrc/sdf.cu), generic (the reference's data flow, calc_color on the kernels), composed (tensor expressions).
Prints one JSON object; recorded under profiles/."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import human_body_reconstruction_b200 as h
from conftest import load_golden
import test_gpu_sdf as T

dev = "cuda"
out = {}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


# ---- fixture errors per route ----------------------------------------------------------------------------
g = load_golden("sdf.npz")
for route in T.ROUTES:
    h.helper.SDF_KERNELS = route != "composed"
    _, enc, mlp, var, vr = T._build(g)
    vr.sdf_native = route == "native"
    S = g["t"].shape[0]
    Cr, _, norm = vr.vol_render(torch.nn.DataParallel(mlp, device_ids=[0]), g["rays_d"].to(dev), g["rays_o"].to(dev), num_samples=S,
                                t=g["t"].to(dev), update_mask=False, dir_norm=g["dir_norm"].to(dev), hierarchical=False)
    gt = g["gt"].to(dev)
    loss = 2 * torch.nn.functional.mse_loss(Cr, gt) + 0.1 * h.helper.eikonal_loss(norm)
    loss.backward()
    rec = {"Cr_abs_over_max": float((Cr.detach().cpu() - g["Cr"]).abs().max() / g["Cr"].abs().max()),
           "norm_abs": float((norm.detach().cpu() - g["norm"]).abs().max()), "loss_rel": abs(float(loss) - float(g["loss"])) / float(g["loss"]),
           "grad_b_rel": abs(float(var.b.grad) - float(g["grad_b"])) / abs(float(g["grad_b"])),
           "dtables_rel": rel(torch.stack([e.weight.grad for e in enc.Embedding_list]), g["dtables"]),
           "mlp_grad_rel_max": max(rel(v.grad, g["grad__" + k.replace(".", "__")]) for k, v in mlp.named_parameters())}
    if route == "native":
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            Cr16, _, _ = vr.vol_render(mlp, g["rays_d"].to(dev), g["rays_o"].to(dev), num_samples=S, t=g["t"].to(dev), update_mask=False,
                                       dir_norm=g["dir_norm"].to(dev), hierarchical=False)
        rec["fp16_autocast_Cr_rel"] = rel(Cr16, Cr.detach())
    out["fixture_" + route] = rec
h.helper.SDF_KERNELS = True

# ---- step time at the C2 shape ---------------------------------------------------------------------------------
MU, MAXB = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
SIGMA = ((MAXB - MU) ** 2).sum().sqrt()
torch.manual_seed(0)
enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=2 ** 19, dim=3, mu=MU.to(dev), sigma=SIGMA.to(dev))
with torch.no_grad():
    for e in enc.Embedding_list:
        e.weight.mul_(1e4)
mlp = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, use_sdf=True, max_bound=MAXB, min_bound=MU)
enc, mlp = enc.to(dev), mlp.to(dev)
var = h.helper.VarModel().to(dev)
vr = h.Volume_Renderer(H=800, W=800, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device=dev, Pos_encode=enc,
                       Dir_encode=h.PositionalEncoder(3, 4), max_dim=1024, sigma_val=SIGMA, mu=MU, use_sdf=True, var_model=var)
R, S = 4096, 128
gen = torch.Generator().manual_seed(1)
ro = (torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.05 * torch.randn(R, 3, generator=gen)).to(dev)
rd = torch.nn.functional.normalize(-ro.cpu() + 0.5 * torch.randn(R, 3, generator=gen), dim=-1).to(dev)
dn = (1 + 0.2 * torch.rand(R, 1, generator=gen)).to(dev)
gt = torch.rand(R, 3, generator=gen).to(dev)
params = list(enc.parameters()) + list(mlp.parameters()) + list(var.parameters())


def step(amp):
    for p in params:
        p.grad = None
    with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
        Cr, Cf, norm = vr.vol_render(mlp, rd, ro, num_samples=S, update_mask=False, dir_norm=dn, hierarchical=False)
        loss = torch.nn.functional.mse_loss(Cr, gt) + torch.nn.functional.mse_loss(Cf, gt) + 0.1 * h.helper.eikonal_loss(norm)
    loss.backward()
    return loss


def timed(amp, reps=5):
    for _ in range(2):
        step(amp)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); l = step(amp); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return {"ms_per_step_min": min(ts), "ms_per_step_median": sorted(ts)[len(ts) // 2], "loss": float(l)}


for route in T.ROUTES:
    h.helper.SDF_KERNELS = route != "composed"
    vr.sdf_native = route == "native"
    h._lib.STATS.reset()
    step(False)
    launches = h._lib.STATS.launches
    out["step_" + route] = {"fp32": timed(False), "fp16_autocast": timed(True), "hbr_launches_per_step": launches}
h.helper.SDF_KERNELS = True
vr.sdf_native = True
mlp.sdf_follows_autocast = True
out["step_native_stencil_follows_autocast"] = {"fp16_autocast": timed(True)}
mlp.sdf_follows_autocast = False
# where the time goes (native route, fp32 stencil): per C-ABI entry point, CUDA events around every call
h._lib.STATS.reset()
h._lib.STATS.timing = True
step(True)
torch.cuda.synchronize()
h._lib.STATS.timing = False
out["native_fp16_step_kernels_ms"] = {k: {"calls": c, "mean_ms": m} for k, (c, m) in sorted(h._lib.STATS.summary().items())}
out["workload"] = f"{R} rays x {S} samples, L=16 F=2 T=2^19, use_sdf, fwd+bwd incl. the eikonal term, eager (host-driven), L2 warm"
print(json.dumps(out))
