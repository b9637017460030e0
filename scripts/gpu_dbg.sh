cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python "$@" > gpurun_out/dbg.log 2>&1; echo "rc=$?"; tail -120 gpurun_out/dbg.log
