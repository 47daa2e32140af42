#!/bin/bash
# round-2 GPU call A: the -m gpu suite, the default bench, and the planner team-layout variants (A/B by MPROS_LIB)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu -s ) > gpurun_out/r2a_gpu_suite.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_gpu_suite.log
grep -E "headline batch|mode \(i\)|mode \(ii\)|  query " gpurun_out/r2a_gpu_suite.log | head -40
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r2a_bench_default.log 2> gpurun_out/r2a_bench_default.err; echo "bench rc=$?"
for V in w4 w6; do
  MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200_$V.so timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r2a_bench_$V.log 2> gpurun_out/r2a_bench_$V.err; echo "$V rc=$?"
done
python - <<'PY'
import json
for f in ["default","w4","w6"]:
    try:
        d=json.loads(open("gpurun_out/r2a_bench_%s.log"%f).read().strip().splitlines()[-1])
        print(f, "plans/s %.0f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "single_batch_ms %.1f"%d["config"]["single_batch_ms"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "e2e %.0f"%d["e2e"]["value"], "col %.3g"%d["extra"]["collision_kernel_checks_per_sec"])
    except Exception as e:
        print(f, "ERR", e)
PY
