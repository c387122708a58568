#!/bin/bash
# Runs bench.py for the listed workloads at N = 1 into gpurun_out/r2_bench_<workload>_n1.json (see tools/collect_profiles.py).
# Usage: tools/run_benches.sh "extra bench args" workload...
extra="$1"; shift
mkdir -p gpurun_out
for w in "$@"; do
  t0=$(date +%s)
  python bench.py --workload $w $extra > gpurun_out/r2_bench_${w}_n1.json 2> gpurun_out/r2_bench_${w}_n1.err
  echo "$w: rc=$? $(( $(date +%s) - t0 )) s"
  tail -2 gpurun_out/r2_bench_${w}_n1.err
  python - <<PY
import json
ls = [l for l in open("gpurun_out/r2_bench_${w}_n1.json") if l.startswith("{")]
if ls:
    d = json.loads(ls[0]); print("$w", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
PY
done
