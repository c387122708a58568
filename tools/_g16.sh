cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest16.log 2>&1; tail -4 gpurun_out/r2_pytest16.log
