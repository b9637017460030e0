# one gpurun call: the SDF kernels (tests + measured errors / step time) and the configs[4] bench leg
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_dropin_scripts.py::test_train_hash2_use_sdf_runs_unmodified_on_the_sdf_kernels -q -x 2>&1 | tail -40 > gpurun_out/pytest_sdf.log ) 2>&1 | grep real; tail -25 gpurun_out/pytest_sdf.log
( time timeout 300 python scripts/measure_sdf.py > gpurun_out/measure_sdf.json 2> gpurun_out/measure_sdf.err ) 2>&1 | grep real; tail -c 3000 gpurun_out/measure_sdf.json; tail -5 gpurun_out/measure_sdf.err
( time timeout 300 python bench.py --no-grid --no-c3 --no-occupancy --no-device-sampler --no-cpu-baseline --steps 10 --repeats 2 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err ) 2>&1 | grep real; tail -3 gpurun_out/bench_c5.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'c5', d.get('c5'))
except Exception as e: print('no bench json', e)
PY
