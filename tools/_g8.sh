cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest8.log 2>&1; tail -5 gpurun_out/r2_pytest8.log
timeout 900 python tools/wbench.py c4s c4n c3b > gpurun_out/r2_wbench8.log 2>&1; cat gpurun_out/r2_wbench8.log
