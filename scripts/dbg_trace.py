import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
from human_body_reconstruction_b200 import _lib
from oracle import port
dev = "cuda"
n = 524288
p = port.mlp_init(seed=5)
flat = torch.cat([v.reshape(-1) for v in p.values()]).to(dev)
feat = (torch.randn(n, 32) * 0.5).to(dev)
dirs = torch.randn(n // 128, 24).to(dev)
out = torch.empty(n, 4, device=dev)
trace = torch.zeros(2048, dtype=torch.int64, device=dev)
L = _lib.lib()
for it in range(3):
    trace.zero_()
    _lib.check(L.hbr_debug_mlp_trace(_lib.ptr(feat), _lib.ptr(dirs), 128, n, _lib.ptr(flat), _lib.ptr(out), _lib.ptr(trace), _lib.stream()))
    torch.cuda.synchronize()
t = trace.cpu().tolist()
g = t[:1000]; m = t[1024:1524]
t0 = g[0]
print("group 0 of CTA 0: per tile 19 stamps: start, then (pre-signal, post-signal, post-wait) x6")
i = 0; tile = 0
while i + 19 <= 1000 and g[i] != 0 and tile < 4:
    s = g[i:i + 19]
    print(f"tile {tile}: start@{s[0]-t0}")
    for k in range(6):
        a, b, c = s[1 + 3 * k], s[2 + 3 * k], s[3 + 3 * k]
        prev = s[0] if k == 0 else s[3 * k]
        mm = m[(tile * 6 + k) * 2:(tile * 6 + k) * 2 + 2]
        print(f"   L{k}: work {a-prev:5d}  signal {b-a:4d}  wait {c-b:5d}   | mma saw ready +{mm[0]-b:5d} after signal, issue+commit {mm[1]-mm[0]:4d}, done seen +{c-mm[1]:5d} after commit")
    i += 19; tile += 1
