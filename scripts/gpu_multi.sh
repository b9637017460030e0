# usage: bash scripts/gpu_multi.sh N   (inside gpurun --gpus N)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5
for NN in 1 $N; do
  if [ $NN -eq 1 ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NN --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 200 $LAUNCH bench.py --gpus $NN --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/multi_${NN}.json 2> gpurun_out/multi_${NN}.err
  echo "N=$NN rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/multi_${NN}.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','eager_ms_per_step','n_gpus')}, d['config']['launch'][:40], 'e2e', d['e2e']['value'])
except Exception as e: print('no json', e); print(open('gpurun_out/multi_${NN}.err').read()[-800:])
PY
done
