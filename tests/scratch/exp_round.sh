#!/bin/bash
# developer run (not a test): full GPU suite + the bench workloads
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 ) > gpurun_out/gpu_suite.log 2>&1; echo "pytest rc=$?"
for W in expand preprocess; do
  timeout 400 python bench.py --workload $W > gpurun_out/bench_$W.log 2> gpurun_out/bench_$W.err; echo "$W rc=$?"
done
timeout 600 python bench.py > gpurun_out/bench_plan_l3.log 2> gpurun_out/bench_plan_l3.err; echo "plan rc=$?"
tail -15 gpurun_out/gpu_suite.log
for f in gpurun_out/bench_expand.log gpurun_out/bench_preprocess.log gpurun_out/bench_plan_l3.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "roof", d["roofline"]["frac"], "cpu", d["cpu_baseline"])
except Exception as e: print("ERR", e); print(open(sys.argv[1].replace(".log",".err")).read()[-1500:])
PY
done
