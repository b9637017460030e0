import os, sys
sys.argv=['x']
exec(open('scripts/bench_hash.py').read().split('# streamed scatter-add')[0])
dy_lm = dy.view(rays * S, 16, 2).permute(1, 0, 2).contiguous()
fn = lambda: ops.hash_encode_bwd_rays_lm(o, d, t, dy_lm, geom, g)
print(f"lm {os.environ.get('HBR_LM_PAIR')}: flushed {timeit(fn, True):7.1f} us   warm {timeit(fn, False):7.1f} us", flush=True)
