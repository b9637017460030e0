cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dropin_scripts.py 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('timed_regions'), 'launches', d['gpu_launches'])
for k,v in d['kernels_ms'].items(): print(f"  {k:24s} {v['mean_ms']*1e3:8.1f} us x{v['launches_per_step']}")
PY
timeout 200 python scripts/dbg_umma_bench.py 2>&1 | tail -40
timeout 200 python scripts/dbg_umma_chain.py 2>&1 | tail -30
