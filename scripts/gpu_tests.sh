cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q "$@" 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
