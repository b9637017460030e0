import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import human_body_reconstruction_b200 as h
from human_body_reconstruction_b200 import _lib
dev = "cuda"
n = 16 * 2 ** 19 * 2
p = torch.randn(n, device=dev); g = torch.randn(n, device=dev) * 1e-3; m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
L = _lib.lib()
def k(): _lib.check(L.hbr_adam_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), n, 0.05, 0.9, 0.999, 1e-8, 0.0, 0, 3, 1.0, None, _lib.stream()))
for _ in range(3): k()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): k()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"adam kernel alone: {ms*1e3:.1f} us  -> {28*n/ms/1e6:.0f} GB/s ({28*n/ms/1e6/6544.7:.2f} of measured HBM peak)")
par = torch.nn.Parameter(p.view(16, -1)); par.grad = g.view(16, -1)
for name, mk in (("hbr FusedAdam", lambda: h.optim.FusedAdam([par], lr=.05)), ("torch fused", lambda: torch.optim.Adam([par], lr=.05, fused=True))):
    o = mk()
    for _ in range(3): o.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20): o.step()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name}: device {e0.elapsed_time(e1)/20*1e3:.1f} us/step, host enqueue {(t1-t0)/20*1e6:.1f} us/step")
