cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final_n2_full.json 2> gpurun_out/final_n2_full.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_n2_full.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['config'].get('allreduce'), d.get('grad_check'), (d.get('c3') or {}).get('value'), (d.get('grid') or {}).get('density_ms'))
PY
