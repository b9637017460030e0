"""(GPU) How far the 16-bit tensor-core MLP is from the reference's fp32 arithmetic -- and how far the REFERENCE'S OWN
16-bit autocast evaluation is from it -- per operand format.  Writes gpurun_out/precision.json (copied to profiles/).

A. fixture size (tests/golden/mlp_autocast.npz: the reference's MLP_3D in fp32, fp16-autocast and bf16-autocast on the CPU)
B. C2 size (4096 rays x 128 samples, L=16 F=2 T=2^19): whole render + loss + backward, fp32 kernels vs fp16 / bf16 operands;
   calibration: oracle.port.mlp_forward (== the reference's torch expressions) under CPU autocast on the same features.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import human_body_reconstruction_b200 as h  # noqa: E402
from human_body_reconstruction_b200 import ops  # noqa: E402
from oracle import port  # noqa: E402
from conftest import load_golden, mlp_params  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def make_mlp(p):
    m = h.MLP_3D(num_sig=2, num_col=2, L=16, F=2, d_view=24, max_bound=torch.ones(3), min_bound=-torch.ones(3))
    m.load_state_dict(p)
    return m.to(DEV)


def fixture_size():
    g = load_golden("mlp_autocast.npz")
    p = mlp_params(g)
    S = int(g["S"])
    res = {}
    for tag, fmt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
        m = make_mlp(p)
        f = g["feat"].to(DEV).requires_grad_()
        out = m.field(f, g["dirs"].to(DEV), S, use_tc=fmt)
        out.backward(g["dout"].to(DEV))
        r = {"out_vs_ref_f32": rel(out, g["f32__out"]), "out_vs_ref_same_fmt": rel(out, g[f"{tag}__out"]),
             "ref_fmt_out_vs_ref_f32": rel(g[f"{tag}__out"], g["f32__out"]),
             "dfeat_vs_ref_f32": rel(f.grad, g["f32__dfeat"]), "dfeat_vs_ref_same_fmt": rel(f.grad, g[f"{tag}__dfeat"]),
             "ref_fmt_dfeat_vs_ref_f32": rel(g[f"{tag}__dfeat"], g["f32__dfeat"])}
        ours, reff, between = {}, {}, {}
        for k, q in m.named_parameters():
            kk = k.replace(".", "__")
            ours[k] = rel(q.grad, g[f"f32__grad__{kk}"])
            between[k] = rel(q.grad, g[f"{tag}__grad__{kk}"])
            reff[k] = rel(g[f"{tag}__grad__{kk}"], g[f"f32__grad__{kk}"])
        r["param_grad_vs_ref_f32_max"] = max(ours.values())
        r["param_grad_vs_ref_same_fmt_max"] = max(between.values())
        r["ref_fmt_param_grad_vs_ref_f32_max"] = max(reff.values())
        r["param_grad_vs_ref_f32"] = ours
        res[tag] = r
    return res


def c2_size(R=4096, S=128, T=2 ** 19):
    torch.manual_seed(0)
    mu, maxb = torch.tensor([-4.27, -4.31, -3.95]), torch.tensor([4.28, 4.27, 2.37])
    sigma = ((maxb - mu) ** 2).sum().sqrt()
    enc = h.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=mu.to(DEV), sigma=sigma.to(DEV))
    with torch.no_grad():
        for e in enc.Embedding_list:
            e.weight.uniform_(-1.0, 1.0)                       # "trained-like" table
    enc = enc.to(DEV)
    p = port.mlp_init(seed=3)
    mlp = make_mlp(p)
    pe = h.PositionalEncoder(3, 4)
    near, far = torch.tensor(2.0), torch.tensor(6.0)
    vr = h.Volume_Renderer(H=800, W=800, K=torch.eye(3), near=near, far=far, device=DEV, Pos_encode=enc, Dir_encode=pe,
                           max_dim=1024, sigma_val=sigma, mu=mu)
    ro = torch.tensor([[0.2, -0.1, 4.0]]).repeat(R, 1) + 0.3 * torch.randn(R, 3)
    rd = torch.nn.functional.normalize(-ro + 0.8 * torch.randn(R, 3), dim=-1)
    gt = torch.rand(R, 3)
    t = port.strat_t(near, far, S, torch.rand(S))
    ro, rd, gt, t = ro.to(DEV), rd.to(DEV), gt.to(DEV), t.to(DEV)

    def run(fmt):
        for q in list(enc.parameters()) + list(mlp.parameters()):
            q.grad = None
        with torch.autocast("cuda", dtype=fmt or torch.bfloat16, enabled=fmt is not None):
            Cr, _, _ = vr.vol_render(mlp, rd, ro, num_samples=S, t=t, dir_norm=1.0, hierarchical=False)
            loss = 2 * torch.nn.functional.mse_loss(Cr, gt)
        loss.backward()
        gtab = torch.stack([e.weight.grad for e in enc.Embedding_list]).clone()
        gm = {k: q.grad.clone() for k, q in mlp.named_parameters()}
        return Cr.detach().clone(), gtab, gm, float(loss)

    base = run(None)
    res = {"R": R, "S": S, "T": T, "loss_f32": base[3]}
    for tag, fmt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
        mlp.tc_grad_scale = 1.0
        Cr, gtab, gm, loss = run(fmt)
        r = {"Cr": rel(Cr, base[0]), "loss": abs(loss - base[3]) / base[3], "table_grad": rel(gtab, base[1]),
             "table_grad_per_level": [rel(gtab[l], base[1][l]) for l in range(16)],
             "mlp_grad_max": max(rel(gm[k], base[2][k]) for k in gm), "mlp_grad": {k: rel(gm[k], base[2][k]) for k in gm}}
        if tag == "f16":
            mlp.tc_grad_scale = 65536.0                       # GradScaler's initial scale (train_hash2.py:156 default)
            Cr2, gtab2, gm2, _ = run(fmt)
            mlp.tc_grad_scale = 1.0
            r["with_grad_scale_65536"] = {"table_grad": rel(gtab2, base[1]), "mlp_grad_max": max(rel(gm2[k], base[2][k]) for k in gm)}
        res[tag] = r
    # calibration: the reference's torch expressions under CPU autocast on the same features / upstream gradient
    with torch.no_grad():
        pts = ops.ray_points(ro, rd, t).view(-1, 3)
        feat = enc(pts).cpu()
    dirs = pe(rd).reshape(R, -1).cpu()
    n = 64 * S                                                 # 64 rays are enough for the statistic; CPU fp16 GEMMs are slow
    drep = dirs[:64, None, :].repeat(1, S, 1).reshape(n, -1)
    dout = torch.randn(n, 4)
    cal = {}
    outs = {}
    for tag, fmt in (("f32", None), ("f16", torch.float16), ("bf16", torch.bfloat16)):
        pa = {k: v.clone().requires_grad_() for k, v in p.items()}
        fa = feat[:n].clone().requires_grad_()
        with torch.autocast("cpu", dtype=fmt or torch.bfloat16, enabled=fmt is not None):
            oa = port.mlp_forward(pa, fa, drep)
        oa.float().backward(dout)
        outs[tag] = (oa.detach().float(), fa.grad.clone(), {k: v.grad.clone() for k, v in pa.items()})
    for tag, fmt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
        f = feat[:n].to(DEV).requires_grad_()
        for q in mlp.parameters():
            q.grad = None
        o = mlp.field(f, dirs[:64].to(DEV), S, use_tc=fmt)
        o.backward(dout.to(DEV))
        cal[tag] = {"ref_autocast_dfeat_vs_f32": rel(outs[tag][1], outs["f32"][1]), "ours_dfeat_vs_f32": rel(f.grad, outs["f32"][1]),
                    "ref_autocast_out_vs_f32": rel(outs[tag][0], outs["f32"][0]), "ours_out_vs_f32": rel(o, outs["f32"][0]),
                    "ref_autocast_param_grad_max": max(rel(outs[tag][2][k], outs["f32"][2][k]) for k in p),
                    "ours_param_grad_max": max(rel(q.grad, outs["f32"][2][k]) for k, q in mlp.named_parameters())}
    res["mlp_on_encoder_features_8192pts"] = cal
    return res


if __name__ == "__main__":
    out = {"fixture_size_1024pts": fixture_size(), "c2_size": c2_size()}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "precision.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
