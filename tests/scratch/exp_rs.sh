#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rs_expand_plan.py -x -q -m gpu -k "device_buffers or rs_" > gpurun_out/rs_test.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/rs_test.log
timeout 400 python bench.py --workload rs > gpurun_out/final_rs.log 2> gpurun_out/final_rs.err; echo "rs rc=$?"
python - gpurun_out/final_rs.log <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"], "cpu", d["cpu_baseline"], d["extra"])
except Exception as e: print("ERR", e); print(open(sys.argv[1].replace(".log",".err")).read()[-2500:])
PY
