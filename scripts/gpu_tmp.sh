cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_gpu_parity.py -m gpu -q -x -s -k "density or grid" 2>&1 | tail -25
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c3 --no-occupancy --no-device-sampler > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d.get('grid'))
PY
