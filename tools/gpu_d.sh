#!/bin/bash
# round-2 GPU call D: duo planner mode (tests + bench A/B), collision streaming-load A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_rs_expand_plan.py tests/test_gpu_synthetic_fullsize.py -q -m gpu -s -k "cluster_mode or 16384 or 4m_poses" ) > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2d_tests.log
grep -E "^FAILED|^ERROR|headline batch" gpurun_out/r2d_tests.log | head
timeout 300 python bench.py --workload collision --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d_col_stream.log 2> gpurun_out/r2d_col_stream.err; echo "stream rc=$?"
MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200_nostream.so timeout 300 python bench.py --workload collision --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d_col_nostream.log 2> gpurun_out/r2d_col_nostream.err; echo "nostream rc=$?"
MPROS_PLAN_MODE=duo timeout 500 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/r2d_bench_duo.log 2> gpurun_out/r2d_bench_duo.err; echo "duo rc=$?"
MPROS_PLAN_MODE=duo timeout 500 python bench.py --steps 4 --warmup 3 --no-cpu --lanes 1 > gpurun_out/r2d_bench_duo_l1.log 2> gpurun_out/r2d_bench_duo_l1.err; echo "duo l1 rc=$?"
python - <<'PY'
import json
for f in ["col_stream","col_nostream"]:
    try:
        d=json.loads(open("gpurun_out/r2d_%s.log"%f).read().strip().splitlines()[-1])
        print(f, "checks/s %.4g"%d["value"], "ms %.4f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/r2d_%s.err"%f).read()[-1500:])
for f in ["bench_duo","bench_duo_l1"]:
    try:
        d=json.loads(open("gpurun_out/r2d_%s.log"%f).read().strip().splitlines()[-1])
        print(f, "plans/s %.0f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "single_batch_ms %.1f"%d["config"]["single_batch_ms"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "e2e %.0f"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/r2d_%s.err"%f).read()[-1500:])
PY
