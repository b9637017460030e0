# usage: bash scripts/gpu_scale.sh   (inside gpurun --gpus 8): weak scaling at the C2 per-GPU batch and at the C3 per-GPU batch
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
run() {  # N rays tag
  NN=$1; RAYS=$2; TAG=$3; shift 3
  if [ $NN -eq 1 ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NN --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 240 $LAUNCH bench.py --gpus $NN --steps 20 --warmup 5 --no-cpu-baseline --rays $RAYS "$@" > gpurun_out/scale_${TAG}_${NN}.json 2> gpurun_out/scale_${TAG}_${NN}.err
  echo "N=$NN rays/gpu=$RAYS rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_${TAG}_${NN}.json').read().strip().splitlines()[-1])
    print('   ', {k:d.get(k) for k in ('value','ms_per_step','eager_ms_per_step','n_gpus')}, d['config']['launch'][:30], 'e2e', round(d['e2e']['value']))
except Exception as e: print('no json', e); print(open('gpurun_out/scale_${TAG}_${NN}.err').read()[-600:])
PY
}
for NN in 1 2 4 8; do run $NN 4096 c2; done
for NN in 1 8; do run $NN 131072 c3 --steps 5 --warmup 3; done
