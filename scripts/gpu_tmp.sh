cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for extra in "" "--no-fuse-scatter"; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-grid --no-occupancy --no-device-sampler $extra > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench $extra rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'c3', d['c3']['value'], d['c3']['ms_per_step'])
except Exception as e: print("no json", e)
PY
done
