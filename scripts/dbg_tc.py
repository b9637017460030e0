import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import port
import human_body_reconstruction_b200 as h
from human_body_reconstruction_b200 import ops
DEV='cuda'
def rel(a,b): a,b=a.detach().double().cpu(),b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
torch.manual_seed(0)
bf=lambda x: x.bfloat16().float()
for mode,(M,N,K) in [(0,(128,64,32)),(0,(128,16,64)),(0,(128,64,48)),(1,(128,64,16)),(1,(128,48,64)),(1,(128,32,64)),(2,(64,64,128)),(2,(64,48,128)),(2,(64,8,128)) if False else (2,(64,32,128))]:
    if mode==0: A=torch.randn(M,K); B=torch.randn(N,K); ref=bf(A)@bf(B).T
    elif mode==1: A=torch.randn(M,K); B=torch.randn(K,N); ref=bf(A)@bf(B)
    else: A=torch.randn(K,64); B=torch.randn(K,N); ref=bf(A).T@bf(B)
    D=ops.debug_umma(mode,A.to(DEV),B.to(DEV),M,N,K)
    torch.cuda.synchronize()
    print('umma mode',mode,(M,N,K),'rel err',rel(D,ref), 'max', (D.cpu()-ref).abs().max().item())
# MLP
p=port.mlp_init(seed=5)
m=h.MLP_3D(num_sig=2,num_col=2,L=16,F=2,d_view=24,max_bound=torch.ones(3),min_bound=-torch.ones(3)); m.load_state_dict(p); m=m.to(DEV)
R,S=40,24
feat=torch.randn(R*S,32)*0.5
dirs=port.dir_encode(torch.nn.functional.normalize(torch.randn(R,3),dim=-1),4)
pr={k:v.clone().requires_grad_() for k,v in p.items()}
fr=feat.clone().requires_grad_()
ref=port.mlp_forward(pr,fr,dirs[:,None,:].repeat(1,S,1).reshape(R*S,-1))
f=feat.to(DEV).requires_grad_()
out=m.field(f,dirs.to(DEV),S,use_tc=True)
torch.cuda.synchronize()
print('mlp tc fwd rel',rel(out,ref),'rgb',rel(out[:,:3],ref[:,:3]),'sigma',rel(out[:,3],ref[:,3]))
out32=m.field(f,dirs.to(DEV),S,use_tc=False)
print('mlp f32 fwd rel',rel(out32,ref))
dout=torch.randn(R*S,4)
ref.backward(dout)
out.backward(dout.to(DEV))
torch.cuda.synchronize()
print('dfeat rel',rel(f.grad,fr.grad))
for k,q in m.named_parameters(): print(' ',k,rel(q.grad,pr[k].grad), float(q.grad.abs().max()))
