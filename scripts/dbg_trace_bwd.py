import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from human_body_reconstruction_b200 import _lib
from oracle import port
dev = "cuda"
n = 524288
p = port.mlp_init(seed=5)
flat = torch.cat([v.reshape(-1) for v in p.values()]).to(dev)
feat = (torch.randn(n, 32) * 0.5).to(dev)
dirs = torch.randn(n // 128, 24).to(dev)
out = torch.randn(n, 4, device=dev)
dout = torch.randn(n, 4, device=dev)
dfeat = torch.empty(n, 32, device=dev)
dparams = torch.zeros_like(flat)
trace = torch.zeros(2048, dtype=torch.int64, device=dev)
L = _lib.lib()
for it in range(3):
    trace.zero_()
    _lib.check(L.hbr_debug_mlp_trace_bwd(_lib.ptr(feat), _lib.ptr(dirs), 128, n, _lib.ptr(flat), _lib.ptr(out), _lib.ptr(dout),
                                         _lib.ptr(dfeat), _lib.ptr(dparams), _lib.ptr(trace), _lib.stream()))
    torch.cuda.synchronize()
from human_body_reconstruction_b200 import ops
from human_body_reconstruction_b200._lib import MlpDims
dims = MlpDims(32, 24)
for nm, fn in (("bwd", lambda: ops.mlp_bwd_tc(feat, dirs, 128, flat, dims, out, dout, True, False, dparams)),
               ("fwd", lambda: ops.mlp_fwd_tc(feat, dirs, 128, flat, dims))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"standalone {nm} kernel (warm L2, incl. output alloc): {e0.elapsed_time(e1)/10*1e3:.1f} us")
t = trace.cpu().tolist()
print("CTA0 clocks: setup", t[2001]-t[2000], "main loop", t[2002]-t[2001], "flush", t[2003]-t[2002])
g = t[:1000]; m = t[1024:1524]
# (the per-stage stamps of earlier kernel versions are gone: the tile groups issue their own GEMMs now)
