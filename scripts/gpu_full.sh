# one gpurun call: smoke, tests, default bench (+ reference arm), launch list and ncu --set full captures (step kernels, grid kernels)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r02d}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --graph off --no-c3 --no-grid --no-occupancy --no-device-sampler"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"hash_fwd_kernel|mlp_fwd_tc_kernel|mlp_bwd_tc_kernel|composite_fwd_kernel|composite_bwd_kernel" -s 15 -c 5 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
GCMD="python scripts/bench_grid.py --res 256"
$GCMD > gpurun_out/${TAG}_grid_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"density_tf32_kernel|hash_fwd_kernel|grid_points_kernel|mc_count_kernel" -s 12 -c 4 -f -o gpurun_out/${TAG}_grid_prof $GCMD > gpurun_out/${TAG}_grid_ncu.log 2>&1
echo "grid capture rc=$?"; tail -2 gpurun_out/${TAG}_grid_plain.log
ls -la gpurun_out | tail -8
