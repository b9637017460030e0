import os, sys
sys.argv=['x']
exec(open('scripts/bench_hash.py').read().split('for name, fn in')[0])
dy_lm = dy.view(rays * S, 16, 2).permute(1, 0, 2).contiguous()
for _ in range(3):
    ops.hash_encode_bwd_rays(o, d, t, dy, geom, g)
    ops.hash_encode_bwd_rays_lm(o, d, t, dy_lm, geom, g)
torch.cuda.synchronize()
