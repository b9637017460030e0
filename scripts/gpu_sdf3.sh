# one gpurun call: SDF tests + measurements (after the density-only backward)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_dropin_scripts.py::test_train_hash2_use_sdf_runs_unmodified_on_the_sdf_kernels -q -x 2>&1 | tail -25 > gpurun_out/pytest_sdf.log ) 2>&1 | grep real; tail -12 gpurun_out/pytest_sdf.log
( time timeout 300 python scripts/measure_sdf.py > gpurun_out/measure_sdf.json 2> gpurun_out/measure_sdf.err ) 2>&1 | grep real; tail -c 2200 gpurun_out/measure_sdf.json; tail -3 gpurun_out/measure_sdf.err
