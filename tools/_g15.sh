cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest15.log 2>&1; tail -4 gpurun_out/r2_pytest15.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/r02_c3_n1_launches.csv python bench.py --steps 32 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02_c3_n1_under_ncu.log 2>&1; tail -c 300 gpurun_out/r02_c3_n1_under_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csx_spmv_kernel -s 4 -c 1 -o gpurun_out/r02_c3_n1_full -f python bench.py --steps 16 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02_c3_n1_full.log 2>&1; tail -2 gpurun_out/r02_c3_n1_full.log
for w in c4s c3b c4n; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csx_spmv_kernel -s 4 -c 1 -o gpurun_out/r02_${w}_full -f python tools/wbench.py $w > gpurun_out/r02_${w}_full.log 2>&1; tail -1 gpurun_out/r02_${w}_full.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csx_stream_kernel -s 4 -c 1 -o gpurun_out/r02_c5s_full -f python tools/wbench.py c5s > gpurun_out/r02_c5s_full.log 2>&1; tail -1 gpurun_out/r02_c5s_full.log
