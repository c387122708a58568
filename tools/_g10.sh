cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest10.log 2>&1; tail -5 gpurun_out/r2_pytest10.log
for w in c3 c4s c5s; do
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 128 --warmup 10 --workload $w > gpurun_out/r2_bench_${w}_n2.json 2> gpurun_out/r2_bench_${w}_n2.err; tail -c 2500 gpurun_out/r2_bench_${w}_n2.json; tail -3 gpurun_out/r2_bench_${w}_n2.err
done
