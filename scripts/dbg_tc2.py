import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import port
import human_body_reconstruction_b200 as h
DEV='cuda'
def rel(a,b): a,b=a.detach().double().cpu(),b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
torch.manual_seed(0)
p=port.mlp_init(seed=5)
m=h.MLP_3D(num_sig=2,num_col=2,L=16,F=2,d_view=24,max_bound=torch.ones(3),min_bound=-torch.ones(3)); m.load_state_dict(p); m=m.to(DEV)
for (R,S) in [(40,24),(3,100),(512,128)]:
    feat=torch.randn(R*S,32)*0.5
    dirs=port.dir_encode(torch.nn.functional.normalize(torch.randn(R,3),dim=-1),4)
    drep=dirs[:,None,:].repeat(1,S,1).reshape(R*S,-1)
    dout=torch.randn(R*S,4)
    pr={k:v.clone().requires_grad_() for k,v in p.items()}
    fr=feat.clone().requires_grad_()
    ref=port.mlp_forward(pr,fr,drep); ref.backward(dout)
    emu_out,emu_dfeat,emu_g,_=port.mlp_bf16_emulation(p,feat,drep,dout)
    f=feat.to(DEV).requires_grad_()
    for q in m.parameters(): q.grad=None
    out=m.field(f,dirs.to(DEV),S,use_tc=True); out.backward(dout.to(DEV)); torch.cuda.synchronize()
    print((R,S),'fwd: gpu-vs-emu',rel(out,emu_out),' gpu-vs-fp32',rel(out,ref),' emu-vs-fp32',rel(emu_out,ref))
    print('  dfeat: gpu-vs-emu',rel(f.grad,emu_dfeat),' gpu-vs-fp32',rel(f.grad,fr.grad),' emu-vs-fp32',rel(emu_dfeat,fr.grad))
    for k,q in m.named_parameters(): print('  ',k,'gpu-vs-emu',rel(q.grad,emu_g[k]),' emu-vs-fp32',rel(emu_g[k],pr[k].grad))
