# closing check: smoke, the test files that cover the code touched last, and a per-kernel view of the configs[4] shape
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-120
( time timeout 400 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_dropin_scripts.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_check.log ) 2>&1 | grep real; cat gpurun_out/pytest_check.log
( time timeout 200 python bench.py --hash-size 22 --samples 256 --hierarchical --steps 5 --warmup 3 --repeats 1 --no-c3 --no-c5 --no-grid --no-occupancy --no-device-sampler --no-cpu-baseline > gpurun_out/r02f_c5_kernels.json 2> gpurun_out/r02f_c5_kernels.err ) 2>&1 | grep real
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02f_c5_kernels.json').read().strip().splitlines()[-1])
    print('ms/step', d['ms_per_step'], 'eager', d['eager_ms_per_step'], d['config']['launch'][:60])
    for k,v in d['kernels_ms'].items(): print(f"  {k:28s} {v['mean_ms']*1e3:9.1f} us x{v['launches_per_step']}")
except Exception as e:
    print('no json', e); print(open('gpurun_out/r02f_c5_kernels.err').read()[-1200:])
PY
