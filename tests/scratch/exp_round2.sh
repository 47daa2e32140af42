#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_case3_arena.py tests/test_gpu_map_collision.py tests/test_gpu_synthetic_fullsize.py tests/test_gpu_rs_expand_plan.py -x -q -m gpu -k "goal or layers or expand" > gpurun_out/goal_test.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/goal_test.log
for W in expand preprocess; do
  timeout 400 python bench.py --workload $W > gpurun_out/bench_$W.log 2> gpurun_out/bench_$W.err; echo "$W rc=$?"
done
for f in gpurun_out/bench_expand.log gpurun_out/bench_preprocess.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "roof", d["roofline"]["frac"], d["extra"])
except Exception as e: print("ERR", e); print(open(sys.argv[1].replace(".log",".err")).read()[-1500:])
PY
done
# ncu: launch list of the preprocess workload, full sets of the two new kernels (after the plain runs above exited 0)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_preprocess.csv python bench.py --workload preprocess --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_pre.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:goal_bfs_cluster -c 1 -o gpurun_out/prof_goal_cluster_r01 -f python bench.py --workload preprocess --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_goal.log 2>&1; echo "ncu goal rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expand_batch -c 1 -o gpurun_out/prof_expand_r01 -f python bench.py --workload expand --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_expand.log 2>&1; echo "ncu expand rc=$?"
