import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from human_body_reconstruction_b200 import _lib
L = _lib.debug_lib()
cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
names = {0: "fwd 128x64x64 (4 MMA)", 1: "dgrad (4 MMA)", 2: "wgrad M64 N72 (8 MMA)", 3: "wgradT M128 N16 (8 MMA)"}
nm = {0: 4, 1: 4, 2: 8, 3: 8}
for kind in range(4):
    for nacc in (1, 4):
        for reps in (1, 8, 64):
            best = None
            for _ in range(3):
                _lib.check(L.hbr_debug_umma_chain_bench(kind, reps, nacc, _lib.ptr(cyc), _lib.stream()))
                torch.cuda.synchronize()
                c = cyc.tolist()
                best = c if best is None or c[0] < best[0] else best
            print(f"{names[kind]:28s} nacc={nacc} reps={reps:3d}: total {best[0]:6d} issue {best[1]:6d}  -> {best[0]/(reps*nm[kind]):6.1f} cyc/MMA")
