#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_map_collision.py tests/test_gpu_synthetic_fullsize.py tests/test_gpu_case3_arena.py tests/test_gpu_collision_methods.py tests/test_gpu_headless_facade.py -x -q -m gpu -k "collision or verdict or layers or facade or headless" > gpurun_out/c1_test.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/c1_test.log
for mode in two walk; do
  if [ $mode = walk ]; then export MPROS_C1_MODE=walk; fi
  timeout 300 python bench.py --workload collision --no-cpu > gpurun_out/bench_c1_$mode.log 2> gpurun_out/bench_c1_$mode.err; echo "$mode rc=$?"
  python - gpurun_out/bench_c1_$mode.log <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["config"]["preprocess_s_first_call"])
except Exception as e: print("ERR", e); print(open(sys.argv[1].replace(".log",".err")).read()[-1500:])
PY
done
