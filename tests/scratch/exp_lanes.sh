#!/bin/bash
# developer experiment (not a test): batches in flight (lanes) and 3 teams per SM
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synthetic_fullsize.py -x -q -m gpu -k "batches_in_flight or full_batch" > gpurun_out/lanes_test.log 2>&1; echo "pytest rc=$?" 
for L in 1 2 3; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --lanes $L > gpurun_out/bench_lanes$L.log 2> gpurun_out/bench_lanes$L.err; echo "lanes $L rc=$?"
done
export MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200_m3.so
for L in 1 3; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --lanes $L > gpurun_out/bench_m3_lanes$L.log 2> gpurun_out/bench_m3_lanes$L.err; echo "m3 lanes $L rc=$?"
done
tail -3 gpurun_out/lanes_test.log
for f in gpurun_out/bench_lanes?.log gpurun_out/bench_m3_lanes?.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("single_batch_ms"), d["gpu_launches"])
except Exception as e: print("ERR", e)
PY
done
