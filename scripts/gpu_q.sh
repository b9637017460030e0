cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "field_bwd_rays" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c3 --no-grid --no-occupancy --no-device-sampler "$@" > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_q.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'step_roofline', d['step_roofline']['frac'])
    for k,v in d['kernels_ms'].items(): print(f"  {k:24s} {v['mean_ms']*1e3:8.1f} us x{v['launches_per_step']}")
except Exception as e: print('no bench json', e)
PY
