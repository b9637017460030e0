cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_peer.py -m gpu -q -x 2>&1 | tail -5
run() {
  TAG=$1; shift
  LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
  timeout 300 $LAUNCH bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/multi_${TAG}.json 2> gpurun_out/multi_${TAG}.err
  echo "N=$N $TAG rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/multi_${TAG}.json').read().strip().splitlines()[-1])
    print('  ', {k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, d['config'].get('allreduce'))
    for k in ('c3','grad_check','grid'):
        if d.get(k): print('  ', k, {a:b for a,b in d[k].items() if a in ('value','ms_per_step','ok','density_ms','mc_count_ms','frac','error','identical','max_rel_err','vertices')})
except Exception as e: print('no json', e); print(open('gpurun_out/multi_${TAG}.err').read()[-1500:])
PY
}
run full
run c2 --no-c3 --no-grid --peer-chunks 2
run c4 --no-c3 --no-grid --peer-chunks 4,8,12,14
run c3b --no-c3 --no-grid --peer-chunks 6,11,14
run c1 --no-c3 --no-grid --peer-chunks 0
