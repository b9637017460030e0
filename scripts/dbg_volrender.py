import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import load_golden, mlp_params
from oracle import port
import human_body_reconstruction_b200 as h
from test_gpu_parity import make_encoder, make_mlp
DEV='cuda'
def rel(a,b): a,b=a.detach().double().cpu(),b.detach().double().cpu(); return float((a-b).norm()/(b.norm()+1e-30))
g=load_golden("volrender.npz")
enc=make_encoder(g); mlp=make_mlp(mlp_params(g,"mlp__"))
S=24; R=g["rays_o"].shape[0]
t=port.strat_t(g["near"],g["far"],S,g["coarse__u_t"])
# oracle stage by stage (fp32 CPU, autograd)
p={k:v.clone().requires_grad_() for k,v in mlp_params(g,"mlp__").items()}
tables=g["tables"].clone().requires_grad_()
pts=port.ray_points(g["rays_o"],g["rays_d"],t).reshape(-1,3)
dirs_r=port.dir_encode(g["rays_d"],4)
feat=port.hash_encode(pts,tables,g["mu"],g["sigma"],g["scales"]); feat.retain_grad()
out=port.mlp_forward(p,feat,dirs_r[:,None,:].repeat(1,S,1).reshape(R*S,-1)); out.retain_grad()
C,w=port.composite(t,out[:,0:3].reshape(R,S,3),out[:,3].reshape(R,S),g["dir_norm"])
loss=2*torch.nn.functional.mse_loss(C,g["gt"]); loss.backward()
# gpu stage by stage
ptsg=h.ops.ray_points(g["rays_o"].to(DEV),g["rays_d"].to(DEV),t.to(DEV)).view(-1,3)
print('pts equal',torch.equal(ptsg.cpu(),pts))
featg=enc(ptsg); featg.retain_grad()
print('feat rel',rel(featg,feat))
dg=h.ops.dir_encode(g["rays_d"].to(DEV),4); print('dir rel',rel(dg,dirs_r))
outg=mlp.field(featg,dg,S,use_tc=False); outg.retain_grad()
print('out rel',rel(outg,out), 'sigma range', out[:,3].min().item(), out[:,3].max().item())
Cg,wg=h.ops.CompositePacked.apply(outg,t.to(DEV),g["dir_norm"].to(DEV),None,R,S)
print('C rel',rel(Cg,C),'w rel',rel(wg,w))
lossg=2*torch.nn.functional.mse_loss(Cg,g["gt"].to(DEV)); lossg.backward()
print('loss',lossg.item(),loss.item())
print('dout rel',rel(outg.grad,out.grad))
print('dfeat rel',rel(featg.grad,feat.grad))
gt_=torch.stack([e.weight.grad for e in enc.Embedding_list])
print('dtables rel',rel(gt_,tables.grad),'vs golden',rel(gt_,g["coarse__dtables"]))
for k,q in mlp.named_parameters(): print(' ',k,rel(q.grad,p[k].grad))
# isolate: feed oracle's dout into gpu mlp bwd
featg2=feat.detach().to(DEV).requires_grad_()
outg2=mlp.field(featg2,dg,S,use_tc=False)
for q in mlp.parameters(): q.grad=None
outg2.backward(out.grad.to(DEV))
print('isolated mlp bwd: dfeat rel',rel(featg2.grad,feat.grad))
for k,q in mlp.named_parameters(): print(' ',k,rel(q.grad,p[k].grad))
# isolate hash bwd with oracle dfeat
for e in enc.Embedding_list: e.weight.grad=None
enc(ptsg).backward(feat.grad.to(DEV))
print('isolated hash bwd rel',rel(torch.stack([e.weight.grad for e in enc.Embedding_list]),tables.grad))
