# usage: bash scripts/gpu_peer.sh N   (inside gpurun --gpus N)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-2}
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 240 $LAUNCH scripts/dbg_peer.py > gpurun_out/peer_${N}.json 2> gpurun_out/peer_${N}.err; echo "dbg_peer rc=$?"
cat gpurun_out/peer_${N}.json; tail -5 gpurun_out/peer_${N}.err
for AR in nccl peer; do
  timeout 200 $LAUNCH bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --allreduce $AR > gpurun_out/peer_bench_${AR}_${N}.json 2> gpurun_out/peer_bench_${AR}_${N}.err
  echo "bench $AR rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/peer_bench_${AR}_${N}.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','eager_ms_per_step','n_gpus','error','last_loss')}, d['config'].get('allreduce'), 'e2e', d['e2e']['value'])
except Exception as e: print('no json', e); print(open('gpurun_out/peer_bench_${AR}_${N}.err').read()[-1500:])
PY
done
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5
