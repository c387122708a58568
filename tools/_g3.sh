cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; tail -3 gpurun_out/r2_pytest3.log
timeout 900 python tools/wbench.py c3b c4n c5s > gpurun_out/r2_wbench3.log 2>&1; cat gpurun_out/r2_wbench3.log
for w in c5s c3b; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csx_stream_kernel -s 3 -c 1 -o gpurun_out/r2_${w}_stream_v3 -f python tools/wbench.py $w > gpurun_out/r2_ncu_$w.log 2>&1; tail -1 gpurun_out/r2_ncu_$w.log
done
