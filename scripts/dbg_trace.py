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
L = _lib.debug_lib()
for it in range(3):
    trace.zero_()
    _lib.check(L.hbr_debug_mlp_trace(_lib.ptr(feat), _lib.ptr(dirs), 128, n, _lib.ptr(flat), _lib.ptr(out), _lib.ptr(trace), _lib.stream()))
    torch.cuda.synchronize()
t = trace.cpu().tolist()
g = [x for x in t[:1000] if x]
# stamps per tile: start, wait_all, convert, then per layer (pre-barrier, post-barrier, post-wait) x 6  = 21
NS = 21
names = ["L0", "L1", "L2", "L3", "L4", "L5"]
m = t[1024:1524]
for tile in range(min(4, len(g) // NS)):
    s = g[tile * NS:(tile + 1) * NS]
    nxt = g[(tile + 1) * NS] - s[0] if len(g) > (tile + 1) * NS else 0
    print(f"tile {tile}: total {nxt}  | features: wait_all {s[1]-s[0]} convert {s[2]-s[1]}")
    prev = s[2]
    for k in range(6):
        a, b, c = s[3 + 3 * k], s[4 + 3 * k], s[5 + 3 * k]
        mm = m[(tile * 6 + k) * 2:(tile * 6 + k) * 2 + 2]
        print(f"   {names[k]}: work {a-prev:5d}  barrier {b-a:4d}  issue {mm[1]-mm[0]:4d}  mma+wake {c-mm[1]:5d}   (stage total {c-prev})")
        prev = c
