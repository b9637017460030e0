import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from human_body_reconstruction_b200 import _lib
L = _lib.debug_lib()
cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
def run(M, N, reps, nacc, mn):
    best = None
    for _ in range(3):
        _lib.check(L.hbr_debug_umma_bench(M, N, reps, nacc, mn, _lib.ptr(cyc), _lib.stream()))
        torch.cuda.synchronize()
        c = cyc.tolist()
        best = c if best is None or c[0] < best[0] else best
    return best
print("M N reps nacc mn | total  issue | per-mma")
for mn in (0, 1):
    for M in (128, 64):
        for N in (16, 64, 128, 256):
            for nacc in (1, 2, 4):
                if nacc * N > 512: continue
                for reps in (1, 4, 8, 32):
                    t = run(M, N, reps, nacc, mn)
                    print(f"{M:3d} {N:3d} {reps:3d} {nacc:2d} {mn} | {t[0]:6d} {t[1]:6d} | {t[0]/reps:7.1f}")
