"""(GPU) hash-grid kernels alone at the bench's C2 shape: CUDA-event time of hbr_hash_encode_fwd_rays / _bwd_rays with the L2
flushed before every launch (as bench.py does between steps) and warm.  Usage: python scripts/bench_hash.py [rays]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
import human_body_reconstruction_b200 as hbr
from human_body_reconstruction_b200 import ops, _lib

dev = "cuda"
rays = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S, T = 128, 2 ** 19
H = W = 800
c2w, K = B.make_cameras(100, 0), B.intrinsics(H, W)
mx, mn = B.scene_bbox(c2w, K, H, W, 2.0, 6.0)
sigma = ((mx - mn) ** 2).sum().sqrt()
enc = hbr.HashEncoder(N_min=16, N_max=2048.0, L=16, F=2, T=T, dim=3, mu=mn.to(dev), sigma=sigma.to(dev)).to(dev)
with torch.no_grad():
    for e in enc.Embedding_list:
        e.weight.mul_(1e4)
o, d, n, gt = (t.to(dev) for t in B.make_batches(c2w, K, H, W, rays, 1, 100)[0])
t = hbr.helper.strat_sampler(torch.tensor(2.0), torch.tensor(6.0), S, device=dev)
geom, table = enc._geom(), enc._flat_table()
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
dy = torch.randn(rays * S, 32, device=dev)
g = torch.zeros_like(table)


def timeit(fn, flush_l2, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if flush_l2:
            flush.zero_()
            g.zero_()                      # the step zero-fills the gradient behind the flush: it sits in L2 when the scatter starts
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for name, fn in (("fwd f32 ", lambda: ops.hash_encode_fwd_rays(o, d, t, table, geom, _lib.HBR_F32)),
                 ("fwd bf16", lambda: ops.hash_encode_fwd_rays(o, d, t, table, geom, _lib.HBR_BF16)),
                 ("bwd     ", lambda: ops.hash_encode_bwd_rays(o, d, t, dy, geom, g))):
    print(f"tile={os.environ.get('HBR_EXTRA_NVCC', 'default'):24s} rays={rays} {name}: flushed {timeit(fn, True):7.1f} us   warm {timeit(fn, False):7.1f} us", flush=True)

# streamed scatter-add (multi-GPU producer): one launch finishing k level chunks in order + per-chunk CTA counts
done = torch.zeros(64, dtype=torch.int32, device=dev)
for k in (1, 2, 4, 8, 16):
    step = 16 // k
    chunks = [(l0, l0 + step) for l0 in range(0, 16, step)]
    if k == 1:
        continue
    fn = lambda: (done.zero_(), ops.hash_encode_bwd_rays_stream(o, d, t, dy, geom, g, chunks, done))
    print(f"bwd streamed, {k:2d} chunks: flushed {timeit(fn, True):7.1f} us   warm {timeit(fn, False):7.1f} us", flush=True)
    launches = lambda: [ops.hash_encode_bwd_rays(o, d, t, dy, geom, g, l0, l1) for l0, l1 in chunks]
    print(f"bwd {k:2d} launches       : flushed {timeit(launches, True):7.1f} us   warm {timeit(launches, False):7.1f} us", flush=True)

# level-major scatter-add (co-resident grid walking the levels in order; d(features) as (L, N, F))
dy_lm = dy.view(rays * S, 16, 2).permute(1, 0, 2).contiguous()
fn = lambda: ops.hash_encode_bwd_rays_lm(o, d, t, dy_lm, geom, g)
print(f"bwd level-major, no counters: flushed {timeit(fn, True):7.1f} us   warm {timeit(fn, False):7.1f} us", flush=True)
for k in (2, 4, 8, 16):
    step = 16 // k
    chunks = [(l0, l0 + step) for l0 in range(0, 16, step)]
    fn = lambda: (done.zero_(), ops.hash_encode_bwd_rays_lm(o, d, t, dy_lm, geom, g, chunks, done))
    print(f"bwd level-major, {k:2d} chunks + counters: flushed {timeit(fn, True):7.1f} us   warm {timeit(fn, False):7.1f} us", flush=True)
