cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for TILE in 128 64 32; do
  export HBR_EXTRA_NVCC="-DHBR_HASH_TILE=$TILE"
  touch human_body_reconstruction_b200/csrc/hash_grid.cu
  python -c "from human_body_reconstruction_b200 import _lib; _lib.build()" > /dev/null 2>&1
  python scripts/bench_hash.py 4096 2>&1 | tail -3
  python scripts/bench_hash.py 131072 2>&1 | tail -3
done
