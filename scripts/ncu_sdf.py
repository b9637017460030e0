"""One native SDF training step at the bench shape (4096 rays x 128 samples, fp16 autocast) for an `ncu --set full` capture
of the csrc/sdf.cu kernels: ncu -k regex:"sdf" ... python scripts/ncu_sdf.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import human_body_reconstruction_b200 as h
dev = "cuda"
MU, MAXB = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
SIGMA = ((MAXB - MU) ** 2).sum().sqrt()
torch.manual_seed(0)
enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=2 ** 19, dim=3, mu=MU.to(dev), sigma=SIGMA.to(dev))
with torch.no_grad():
    for e in enc.Embedding_list:
        e.weight.mul_(1e4)
mlp = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=MAXB, min_bound=MU)      # as train_hash2.py:127 builds it
enc, mlp = enc.to(dev), mlp.to(dev)
var = h.helper.VarModel().to(dev)
vr = h.Volume_Renderer(H=800, W=800, K=torch.eye(3), near=torch.tensor(2.0), far=torch.tensor(6.0), device=dev, Pos_encode=enc,
                       Dir_encode=h.PositionalEncoder(3, 4), max_dim=1024, sigma_val=SIGMA, mu=MU, use_sdf=True, var_model=var)
R, S = 4096, 128
gen = torch.Generator().manual_seed(1)
ro = (torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.05 * torch.randn(R, 3, generator=gen)).to(dev)
rd = torch.nn.functional.normalize(-ro.cpu() + 0.5 * torch.randn(R, 3, generator=gen), dim=-1).to(dev)
gt = torch.rand(R, 3, generator=gen).to(dev)
for _ in range(2):
    with torch.autocast("cuda", dtype=torch.float16):
        Cr, Cf, norm = vr.vol_render(mlp, rd, ro, num_samples=S, update_mask=False, dir_norm=1.0, hierarchical=False)
        loss = 2 * torch.nn.functional.mse_loss(Cr, gt) + 0.1 * h.helper.eikonal_loss(norm)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
