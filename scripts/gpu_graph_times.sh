# one short gpurun call: default step only, per-kernel times of the eager region against the in-graph external events
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 200 python bench.py --no-c3 --no-grid --no-occupancy --no-device-sampler --no-c5 --no-cpu-baseline > gpurun_out/r02f_graph_times.json 2> gpurun_out/r02f_graph_times.err ) 2>&1 | grep real; tail -3 gpurun_out/r02f_graph_times.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02f_graph_times.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'roofline', {k:d['roofline'][k] for k in ('kernel','frac','mean_launch_ms','timing')})
    g=d['kernels_ms_in_graph']
    for k,v in d['kernels_ms'].items():
        gg=g.get(k)
        print(f"  {k:28s} eager {v['mean_ms']*1e3:8.1f} us x{v['launches_per_step']}   in graph {gg['mean_ms']*1e3 if gg else float('nan'):8.1f} us x{gg['launches_per_step'] if gg else 0}")
    if 'unavailable' in g: print(g)
    print({k:(round(v['frac'],3)) for k,v in d['kernel_rooflines'].items()})
except Exception as e:
    print('no json', e)
PY
