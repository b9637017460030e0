cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-8}
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 200 $LAUNCH scripts/dbg_peer.py > gpurun_out/peer_${N}.json 2> gpurun_out/peer_${N}.err; echo "dbg_peer rc=$?"
cat gpurun_out/peer_${N}.json; grep -v Warning gpurun_out/peer_${N}.err | tail -3
bash scripts/gpu_peer2.sh $N "--allreduce nccl" "--allreduce peer --peer-ctas 96" "--allreduce peer --peer-transport symm --peer-ctas 64"
