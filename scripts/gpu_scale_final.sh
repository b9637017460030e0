# usage (inside gpurun --gpus 8): bash scripts/gpu_scale_final.sh  -- chunk-list A/B at 8 ranks, then the full bench line with the
# best list, then short lines at 4 and 2 ranks (the 1-rank line comes from the single-GPU run)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
SHORT="--steps 20 --warmup 5 --repeats 3 --no-cpu-baseline --no-c3 --no-grid --no-occupancy --no-device-sampler"
launch() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511"; }
show() {  # file label
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    gc=d.get('grad_check') or {}
    c3=d.get('c3') or {}
    print(f"[{sys.argv[2]:44s}] N={d['n_gpus']} ms={d['ms_per_step']:.4f} value={d['value']/1e6:.2f}M ok={gc.get('ok')} ident={gc.get('ranks_bit_identical')} c3={c3.get('value',0)/1e6:.1f}M grid={(d.get('grid') or {}).get('density_ms')} err={d.get('error')}")
except Exception as e: print(f"[{sys.argv[2]}] no json", e); print(open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
}
best=""; bestms=999
for C in 4 3 6,10,13 6,9,12,14; do
  f=gpurun_out/final_ab_${C//,/_}.json
  timeout 150 $(launch 8) bench.py --gpus 8 $SHORT --peer-chunks $C > $f 2> ${f%.json}.err
  show $f "N=8 chunks $C"
  ms=$(python -c "import json,sys; print(json.loads(open('$f').read().strip().splitlines()[-1])['ms_per_step'])" 2>/dev/null || echo 999)
  if python -c "import sys; sys.exit(0 if float('$ms') < float('$bestms') else 1)"; then best=$C; bestms=$ms; fi
done
echo "best chunk list: $best ($bestms ms)"
timeout 300 $(launch 8) bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --peer-chunks $best > gpurun_out/final_n8_full.json 2> gpurun_out/final_n8_full.err
show gpurun_out/final_n8_full.json "N=8 full, chunks $best"
for N in 4 2; do
  timeout 150 $(launch $N) bench.py --gpus $N $SHORT --peer-chunks $best > gpurun_out/final_n$N.json 2> gpurun_out/final_n$N.err
  show gpurun_out/final_n$N.json "N=$N chunks $best"
done
