# usage: bash scripts/gpu_profile.sh <tag>   (one gpurun call: plain run, launch list, one full capture of the 4 hot kernels)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"hash_fwd_kernel|hash_bwd_kernel|mlp_fwd_tc_kernel|mlp_bwd_tc_kernel|field_" -s 12 -c 4 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -12
tail -3 gpurun_out/${TAG}_ncu_full.log
