#!/bin/bash
mkdir -p gpurun_out
for v in "" _r6 _r8; do
  MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200$v.so timeout 300 python bench.py --workload rs --no-cpu --steps 5 > gpurun_out/bench_rs$v.log 2> gpurun_out/bench_rs$v.err; echo "rs$v rc=$?"
  python - gpurun_out/bench_rs$v.log <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
