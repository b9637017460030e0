# gpurun --gpus 2: SDF mode sharded over 2 ranks (gradient parity) + the N=2 bench line (configs[2] leg on the median step)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k sdf 2>&1 | tail -15 > gpurun_out/pytest_sdf_n2.log ) 2>&1 | grep real; cat gpurun_out/pytest_sdf_n2.log
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-grid > gpurun_out/r02f_n2.json 2> gpurun_out/r02f_n2.err ) 2>&1 | grep real; echo rc=$?
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02f_n2.json').read().strip().splitlines()[-1])
    c3=d.get('c3') or {}
    print(d['ms_per_step'], d['value'], d['config'].get('allreduce'), d.get('grad_check'), c3.get('value'), c3.get('ms_per_step'), c3.get('ms_per_step_mean'))
except Exception as e:
    print('no json', e); print(open('gpurun_out/r02f_n2.err').read()[-1500:])
PY
