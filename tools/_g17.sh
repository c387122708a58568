cd $GRAFT_REPO_ROOT
run() { # workload nproc timeout
(time timeout $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --steps 128 --warmup 10 --workload $1) > gpurun_out/r2_bench_$1_n$2.json 2> gpurun_out/r2_bench_$1_n$2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_bench_$1_n$2.json') if l.startswith('{')][0])
    print("$1 N=$2", round(d['value'],1),'GFLOP/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],3), 'konly', d['roofline']['kernel_only_ms'], 'e2e',round(d['e2e']['value'],1), d['detail']['encoding_rank0'], d['detail']['checks_vs_csr'], 'tune',d['detail']['tune_s'],'gen',d['detail']['generate_s'])
except Exception as e: print("$1 N=$2 failed", e)
PY
grep -E "real|Error|error|failed|timed out" gpurun_out/r2_bench_$1_n$2.err | tail -4
}
run c5 8 900
run c3 8 600
run c3 4 600
run c5 4 900
run c4 4 600
