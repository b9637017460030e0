import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import load_golden
from oracle import port
import human_body_reconstruction_b200 as h
DEV='cuda'
def rel(a,b): a,b=a.double().cpu(),b.double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
for name in ["composite.npz","composite_perray.npz"]:
    g=load_golden(name)
    rgb=g["rgb"].to(DEV).requires_grad_(); sig=g["sigma"].to(DEV).requires_grad_()
    C,w,_=h.helper.calc_color(t=g["t"].to(DEV),rgb=rgb,sigma=sig,dir_norm=g["dir_norm"].to(DEV))
    print(name,'C rel',rel(C,g["C"]),'max abs',(C.cpu()-g["C"]).abs().max().item(),'w rel',rel(w[...,0],g["w"]), 'w maxabs',(w[...,0].cpu()-g["w"]).abs().max().item())
    bad=~torch.isclose(w[...,0].cpu(),g["w"],rtol=1e-5,atol=1e-7)
    print(' w bad count',bad.sum().item(), 'examples', w[...,0].cpu()[bad][:5], g["w"][bad][:5])
    C.backward(g["gC"].to(DEV))
    d64=port.composite_bwd(*(g[k].double() for k in ("t","rgb","sigma","dir_norm","gC")))
    print(' drgb rel',rel(rgb.grad,g["drgb"]),' dsig rel vs f64',rel(sig.grad,d64[1]),' vs golden',rel(sig.grad,g["dsigma"]), 'golden vs f64', rel(g["dsigma"],d64[1]))
    e=(sig.grad.cpu().double()-d64[1]).abs()
    i=e.argmax(); print(' worst',i.item()//e.shape[1], i.item()%e.shape[1], e.max().item(), d64[1].reshape(-1)[i].item(), d64[1].abs().max().item())
    rowerr=(sig.grad.cpu().double()-d64[1]).norm(dim=1)/d64[1].norm(dim=1); print(' row rel err top', rowerr.topk(5))
