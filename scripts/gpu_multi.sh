# usage: bash scripts/gpu_multi.sh N   (inside gpurun --gpus N)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for G in off on; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --graph $G > gpurun_out/multi_${N}_$G.json 2> gpurun_out/multi_${N}_$G.err
  echo "N=$N graph=$G rc=$?"; tail -3 gpurun_out/multi_${N}_$G.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/multi_${N}_$G.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','eager_ms_per_step','n_gpus','allreduce_bytes_per_step')}, d['config']['launch'][:60], d['e2e']['value'])
except Exception as e: print('no json', e)
PY
done
