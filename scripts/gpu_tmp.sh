cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
run() {
  TAG=$1; NN=$2; shift 2
  if [ $NN -eq 1 ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NN --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 300 $LAUNCH bench.py --gpus $NN --steps 20 --warmup 5 "$@" > gpurun_out/scale_${TAG}.json 2> gpurun_out/scale_${TAG}.err
  echo "N=$NN $TAG rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_${TAG}.json').read().strip().splitlines()[-1])
    print('  ', {k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, d['config'].get('allreduce'))
    for k in ('c3','grad_check','grid'):
        if d.get(k): print('  ', k, {a:b for a,b in d[k].items() if a in ('value','ms_per_step','ok','density_ms','mc_count_ms','frac','error','ranks_bit_identical','rel_err_vs_recomputed_mean_of_local_grads','vertices')})
except Exception as e: print('no json', e); print(open('gpurun_out/scale_${TAG}.err').read()[-1500:])
PY
}
run n1 1 --no-cpu-baseline --no-occupancy --no-device-sampler
run n8_full 8
run n8_c0 8 --no-c3 --no-grid --peer-chunks 0
run n8_c5 8 --no-c3 --no-grid --peer-chunks 4,8,12,14
run n8_c3 8 --no-c3 --no-grid --peer-chunks 8,12
run n4_full 4
