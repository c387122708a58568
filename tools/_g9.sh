cd $GRAFT_REPO_ROOT
nproc; free -g | head -2
(time timeout 600 python bench.py --workload small --steps 32 --warmup 3) > gpurun_out/r2_bench_small.json 2> gpurun_out/r2_bench_small.err; tail -c 1500 gpurun_out/r2_bench_small.json; tail -5 gpurun_out/r2_bench_small.err
(time timeout 1500 python bench.py --steps 128 --warmup 10) > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err; tail -c 3000 gpurun_out/r2_bench_c3_n1.json; tail -5 gpurun_out/r2_bench_c3_n1.err
