cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; tail -3 gpurun_out/r2_pytest4.log
timeout 900 python tools/wbench.py c3b c4n c5s > gpurun_out/r2_wbench4.log 2>&1; cat gpurun_out/r2_wbench4.log
