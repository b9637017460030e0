# one gpurun call: smoke, gpu tests, default bench line (what the driver runs at round end)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ) 2>&1 | grep real; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-300
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log ) 2>&1 | grep real; cat gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline'], 'step_roofline', d.get('step_roofline'))
    for k,v in d['kernels_ms'].items(): print(f"  {k:24s} {v['mean_ms']*1e3:8.1f} us x{v['launches_per_step']}")
except Exception as e: print('no bench json', e)
PY
