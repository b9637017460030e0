# usage: bash scripts/gpu_ncu_one.sh <tag> <kernel regex> [skip]   (one full ncu capture of one launch of one kernel)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=$1; RX=$2; SKIP=${3:-3}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --graph off"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c 1 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
