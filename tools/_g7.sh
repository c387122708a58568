cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest7.log 2>&1; tail -5 gpurun_out/r2_pytest7.log
timeout 900 python tools/wbench.py c4s c4n c3b c2s > gpurun_out/r2_wbench7.log 2>&1; cat gpurun_out/r2_wbench7.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csx_ -s 3 -c 1 -o gpurun_out/r2_c4s_v3 -f python tools/wbench.py c4s > gpurun_out/r2_ncu_c4s.log 2>&1; tail -1 gpurun_out/r2_ncu_c4s.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csx_ -s 3 -c 1 -o gpurun_out/r2_c3b_v5 -f python tools/wbench.py c3b > gpurun_out/r2_ncu_c3b.log 2>&1; tail -1 gpurun_out/r2_ncu_c3b.log
