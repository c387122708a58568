cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest11.log 2>&1; tail -4 gpurun_out/r2_pytest11.log
for w in c4 c3b c2; do
(time timeout 1500 python bench.py --steps 128 --warmup 10 --workload $w) > gpurun_out/r2_bench_${w}_n1.json 2> gpurun_out/r2_bench_${w}_n1.err; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_${w}_n1.json') if l.startswith('{')][0])
print("$w", round(d['value'],1),'GFLOP/s', round(d['ms_per_step'],4),'ms frac',round(d['roofline']['frac'],3), 'e2e',round(d['e2e']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'), d['detail']['encoding_rank0'], d['detail']['checks_vs_csr'], 'tune',d['detail']['tune_s'],'gen',d['detail']['generate_s'], d['roofline']['bytes'])
PY
tail -4 gpurun_out/r2_bench_${w}_n1.err
done
