# usage: bash scripts/gpu_peer2.sh N "<extra bench args A>" "<extra bench args B>" ...   (inside gpurun --gpus N)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=$1; shift
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
i=0
for EXTRA in "$@"; do
  i=$((i+1))
  timeout 200 $LAUNCH bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $EXTRA > gpurun_out/peer2_${N}_$i.json 2> gpurun_out/peer2_${N}_$i.err
  echo "[$EXTRA] rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/peer2_${N}_$i.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','eager_ms_per_step','n_gpus','error','last_loss')}, d['config'].get('allreduce'), 'e2e', d['e2e']['value'])
except Exception as e: print('no json', e); print(open('gpurun_out/peer2_${N}_$i.err').read()[-1500:])
PY
done
