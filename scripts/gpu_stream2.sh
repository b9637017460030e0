# usage (inside gpurun --gpus N): bash scripts/gpu_stream2.sh N  -- streamed gradient exchange: tests, then bench A/B at N ranks
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -8
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
i=0
for ARGS in "--peer-exchange stream" "--peer-exchange launch" "--peer-exchange stream --peer-chunks 4" "--peer-exchange stream --peer-chunks 16" "--peer-exchange stream --peer-ctas 64" "${@:2}"; do
  [ -z "$ARGS" ] && continue
  i=$((i+1))
  timeout 200 $LAUNCH bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-c3 --no-grid --no-occupancy --no-device-sampler $ARGS > gpurun_out/stream_${N}_$i.json 2> gpurun_out/stream_${N}_$i.err
  echo "[$ARGS] rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/stream_${N}_$i.json').read().strip().splitlines()[-1])
    print('   ', {k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, d['config'].get('allreduce'), 'grad_check', d.get('grad_check'), d.get('error'))
except Exception as e: print('no json', e); print(open('gpurun_out/stream_${N}_$i.err').read()[-1500:])
PY
done
