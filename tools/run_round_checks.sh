#!/bin/bash
# developer run (not a test): what the driver runs at round end + the bench workloads + one ncu capture
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/gpu_suite.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
for W in plan expand preprocess collision rs; do
  timeout 600 python bench.py --workload $W > gpurun_out/final_$W.log 2> gpurun_out/final_$W.err; echo "$W rc=$?"
  python - gpurun_out/final_$W.log <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
except Exception as e: print("ERR", e); print(open(sys.argv[1].replace(".log",".err")).read()[-1500:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:expand_batch -c 1 -o gpurun_out/prof_expand_r01 -f python bench.py --workload expand --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_expand.log 2>&1; echo "ncu expand rc=$?"
