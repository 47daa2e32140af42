#!/bin/bash
# round-2 GPU call B: banded map with the in-process communicator, the 8-shots-per-warp Reeds-Shepp kernel (tests, bench,
# register-cap variants, ncu), planner team-layout variants
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_banded_local.py tests/test_gpu_rs_expand_plan.py tests/test_gpu_headless_facade.py tests/test_gpu_map_collision.py -x -q -m gpu -s ) > $O/r2b_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2b_tests.log
grep -E "rs_solve|rs_shot|FAILED|Error|error" $O/r2b_tests.log | head -20
timeout 300 python bench.py --workload rs --steps 20 --warmup 5 > $O/r2b_rs.json 2> $O/r2b_rs.err; echo "rs rc=$?"
for V in rs6 rs4; do
  MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200_$V.so timeout 300 python bench.py --workload rs --steps 20 --warmup 5 --no-cpu > $O/r2b_$V.json 2> $O/r2b_$V.err; echo "$V rc=$?"
done
for V in w6 w4; do
  MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200_$V.so timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu --no-arena > $O/r2b_plan_$V.json 2> $O/r2b_plan_$V.err; echo "$V rc=$?"
done
python - <<'PY'
import json
for f in ["rs","rs6","rs4"]:
    try:
        d=json.loads(open("gpurun_out/r2b_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "shots/s %.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/r2b_%s.err"%f).read()[-1500:])
for f in ["plan_w6","plan_w4"]:
    try:
        d=json.loads(open("gpurun_out/r2b_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "plans/s %.0f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "single_batch_ms %.1f"%d["config"]["single_batch_ms"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "e2e %.0f"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/r2b_%s.err"%f).read()[-1500:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_batch -c 1 -f -o $O/r2b_rs_batch_kernel python bench.py --workload rs --steps 1 --warmup 3 --no-cpu > $O/r2b_ncu_rs.log 2>&1; echo "ncu rc=$?"
ncu -i $O/r2b_rs_batch_kernel.ncu-rep --page raw --csv > $O/r2b_rs_batch_kernel_full.csv 2>/dev/null
