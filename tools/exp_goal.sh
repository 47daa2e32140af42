#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_case3_arena.py tests/test_gpu_map_collision.py tests/test_gpu_synthetic_fullsize.py -x -q -m gpu -k "goal" > gpurun_out/goal_test.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/goal_test.log
for mode in default cluster8; do
  if [ $mode = cluster8 ]; then export MPROS_GOAL_MODE=cluster8; fi
  timeout 200 python bench.py --workload preprocess --no-cpu --steps 5 > gpurun_out/bench_pre_$mode.log 2> gpurun_out/bench_pre_$mode.err; echo "$mode rc=$?"
  python - gpurun_out/bench_pre_$mode.log <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["extra"]["update_goal_ms"], d["extra"]["set_occupancy_ms"])
except Exception as e: print("ERR", e); print(open(sys.argv[1].replace(".log",".err")).read()[-800:])
PY
done
