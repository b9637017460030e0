"""CPU analysis (no GPU): how many red.global operations hash_bwd_kernel issues per level on the bench batch after merging runs
of consecutive lanes in the same cell and pairing (x, x+1) corners of an even x.  python scripts/analysis_red_count.py"""
import sys, math, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import port
torch.manual_seed(0)
H=W=800
c2w, K = bench.make_cameras(100,0), bench.intrinsics(H,W)
mx, mn = bench.scene_bbox(c2w,K,H,W,2.0,6.0)
sigma = ((mx-mn)**2).sum().sqrt()
o,d,n,gt = bench.make_batches(c2w,K,H,W,4096,1,100)[0]
S=128
t = torch.linspace(2.0,6.0,S) + torch.rand(S)*4.0/S
pts = (o[:,None,:] + d[:,None,:]*t[None,:,None]).reshape(-1,3)
scales = port.level_scales(16, 2048.0, 16)
T=2**19
tot_runs=0; tot_ops=0
print("level scale  cells/dim  runs/warp  REDs(v2+v4 pairing)  distinct-entries")
for l in range(16):
    s = float(scales[l])
    xs = ((pts - mn)/sigma)*s
    cell = xs.floor().long()
    key = (cell[:,0]*4096 + cell[:,1])*4096 + cell[:,2]
    kw = key.view(-1,32)                      # warps of 32 consecutive samples of a ray
    newrun = torch.ones_like(kw, dtype=torch.bool); newrun[:,1:] = kw[:,1:] != kw[:,:-1]
    runs = int(newrun.sum())
    even = (cell[:,0] % 2 == 0).view(-1,32)
    ops = int((newrun & even).sum())*6 + int((newrun & ~even).sum())*8      # even x: 2 v4 + 4 v2... = 4 pairs -> (pairs along x: 4) -> 4 v4 ; odd: 8 v2
    ops_even = int((newrun & even).sum())*4 + int((newrun & ~even).sum())*8
    tot_runs += runs; tot_ops += ops_even
    print(f"{l:2d} {s:8.1f} {runs/ kw.shape[0]:8.2f} runs/warp   ops {ops_even/1e6:6.2f} M   unique cells {len(torch.unique(key))}")
print("total RED instructions-lanes (M):", tot_ops/1e6, " unmerged would be", 16*8*pts.shape[0]/1e6)
