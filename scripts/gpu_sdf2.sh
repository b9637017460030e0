# one gpurun call: SDF measurements (errors per route, step times, per-kernel times) + ncu --set full of the csrc/sdf.cu kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 300 python scripts/measure_sdf.py > gpurun_out/measure_sdf.json 2> gpurun_out/measure_sdf.err ) 2>&1 | grep real; tail -c 2500 gpurun_out/measure_sdf.json; tail -3 gpurun_out/measure_sdf.err
python scripts/ncu_sdf.py > gpurun_out/ncu_sdf_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sdf" -s 5 -c 5 -f -o gpurun_out/r02f_sdf_prof python scripts/ncu_sdf.py > gpurun_out/ncu_sdf.log 2>&1
echo "sdf capture rc=$?"; tail -2 gpurun_out/ncu_sdf_plain.log; ls -la gpurun_out | tail -5
