set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; tail -5 gpurun_out/r2_pytest1.log
timeout 900 python tools/wbench.py c3b c4n c4s c5s > gpurun_out/r2_wbench1.log 2>&1; cat gpurun_out/r2_wbench1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csx_stream_kernel -s 3 -c 1 -o gpurun_out/r2_c5s_stream_v1 -f python tools/wbench.py c5s > gpurun_out/r2_ncu_c5s.log 2>&1; tail -3 gpurun_out/r2_ncu_c5s.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csx_stream_kernel -s 3 -c 1 -o gpurun_out/r2_c3b_stream_v1 -f python tools/wbench.py c3b > gpurun_out/r2_ncu_c3b.log 2>&1; tail -3 gpurun_out/r2_ncu_c3b.log
