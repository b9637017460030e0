# usage (inside gpurun --gpus N): bash scripts/gpu_stream8.sh N  -- streamed gradient exchange: bench A/B at N ranks
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-8}
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
i=0
while IFS= read -r ARGS; do
  [ -z "$ARGS" ] && continue
  i=$((i+1))
  timeout 150 $LAUNCH bench.py --gpus $N --steps 20 --warmup 5 --repeats 3 --no-cpu-baseline --no-c3 --no-grid --no-occupancy --no-device-sampler $ARGS > gpurun_out/s8_${N}_$i.json 2> gpurun_out/s8_${N}_$i.err
  rc=$?
  python - "$ARGS" $rc <<PY
import json, sys
try:
    d=json.loads(open('gpurun_out/s8_${N}_$i.json').read().strip().splitlines()[-1])
    gc=d.get('grad_check') or {}
    print(f"[{sys.argv[1]:60s}] rc={sys.argv[2]} ms={d['ms_per_step']:.4f} value={d['value']/1e6:.2f}M ok={gc.get('ok')} ident={gc.get('ranks_bit_identical')} err={d.get('error')}")
except Exception as e: print(f"[{sys.argv[1]}] rc={sys.argv[2]} no json", e); print(open('gpurun_out/s8_${N}_$i.err').read()[-800:])
PY
done <<'CFG'
--peer-exchange launch
--peer-exchange stream --peer-scatter lm
--peer-exchange stream --peer-scatter lm --peer-ctas 64
--peer-exchange stream --peer-scatter tile --peer-chunks 8,12,14,15
--peer-exchange stream --peer-scatter tile --peer-chunks 10,13,15
--peer-exchange stream --peer-scatter tile --peer-chunks 4
--peer-exchange stream --peer-scatter tile --peer-chunks 8,12,14,15 --peer-ctas 64
CFG
