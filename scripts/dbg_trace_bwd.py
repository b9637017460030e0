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
L = _lib.debug_lib()
from human_body_reconstruction_b200 import ops
from human_body_reconstruction_b200._lib import MlpDims
scratch = ops.mlp_tc_scratch(MlpDims(32, 24), dev)
for it in range(3):
    trace.zero_()
    _lib.check(L.hbr_debug_mlp_trace_bwd(_lib.ptr(feat), _lib.ptr(dirs), 128, n, _lib.ptr(flat), _lib.ptr(out), _lib.ptr(dout),
                                         _lib.ptr(dfeat), _lib.ptr(dparams), _lib.ptr(scratch), _lib.ptr(trace), _lib.stream()))
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
print("CTA0 clocks (product configuration: operand image + gradient rows): setup", t[2001]-t[2000], "group-0 tile loop", t[2004]-t[2001],
      "tail wait", t[2002]-t[2004], "flush", t[2003]-t[2002])
# per-stage stamps of tile group 0, thread 0 (tiles 2..11 averaged): index 0 = tile start, F0..F4: (entry, done) pairs,
# B5..B0: (entry, bar1, issued, trail done, bar2+arrive, dgrad done), then d(feat) written, tile end
import statistics as st
def avg(i, j):
    return st.mean(t[k * 80 + j] - t[k * 80 + i] for k in range(2, 12))
print(f"tile total {avg(0, 48):.0f}   load+prefetch {avg(0, 1):.0f}")
names = ["F0", "F1", "F2", "F3", "F4"]
for s_, nm in enumerate(names):
    a, b = 1 + 2 * s_, 2 + 2 * s_
    nxt = 3 + 2 * s_ if s_ < 4 else 11
    print(f"  {nm}: barrier+issue+mma+wake {avg(a, b):6.0f}   epilogue after {avg(b, nxt):6.0f}")
for s_, nm in enumerate(["B5", "B4", "B3", "B2", "B1", "B0"]):
    o = 11 + 6 * s_
    nxt = o + 6 if s_ < 5 else 47
    print(f"  {nm}: bar1 {avg(o, o+1):5.0f} issue {avg(o+1, o+2):5.0f} trail {avg(o+2, o+3):5.0f} bar2 {avg(o+3, o+4):5.0f} "
          f"wait dgrad {avg(o+4, o+5):5.0f}   epilogue after {avg(o+5, nxt):6.0f}   (stage {avg(o, nxt):6.0f})")
print(f"  final wait for the last weight-gradient GEMM {avg(47, 48):.0f}")
