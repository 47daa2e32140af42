#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_case3_arena.py tests/test_gpu_map_collision.py tests/test_gpu_synthetic_fullsize.py -x -q -m gpu -k "goal or layers" > gpurun_out/goal_test.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/goal_test.log
timeout 400 python bench.py --workload preprocess --no-cpu > gpurun_out/bench_preprocess.log 2> gpurun_out/bench_preprocess.err; echo "pre rc=$?"
python - gpurun_out/bench_preprocess.log <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "roof", d["roofline"]["frac"], d["extra"])
except Exception as e: print("ERR", e); print(open(sys.argv[1].replace(".log",".err")).read()[-1500:])
PY
