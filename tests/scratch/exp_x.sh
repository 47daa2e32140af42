#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rs_expand_plan.py -x -q -m gpu -k "expand" > gpurun_out/x_test.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/x_test.log
timeout 400 python bench.py --workload expand --no-cpu > gpurun_out/bench_expand.log 2> gpurun_out/bench_expand.err; echo "expand rc=$?"
python - gpurun_out/bench_expand.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
