cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest14.log 2>&1; tail -4 gpurun_out/r2_pytest14.log
timeout 900 python tools/wbench.py c4s c4n c3b c5s 2>&1 | grep rows
