#!/bin/bash
# round-2 final verification run (developer run): GPU suite, smoke, default bench, launch list and full ncu captures of the
# two planner kernels. Tight timeouts: a hang must not eat the GPU budget.
O=gpurun_out; mkdir -p $O
( time timeout 480 python -m pytest tests -x -q -m gpu -s ) > $O/r02f_gpu_suite.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02f_gpu_suite.log
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02f_smoke.log
timeout 600 python bench.py > $O/r02f_bench_plan.json 2> $O/r02f_bench_plan.err; echo "plan rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02f_bench_plan.json").read().strip().splitlines()[-1])
    print("plan", "%.5g"%d["value"], "ms %.4g"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], "single %.1f"%d["config"]["single_batch_ms"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "launches", d["gpu_launches"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, "arena", d["extra"].get("case3_arena",{}).get("plans_per_sec"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02f_bench_plan.err").read()[-1500:])
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -c 200 --csv --log-file $O/r02f_launches_plan_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-arena > $O/r02f_ncu_plan.log 2>&1; echo "launches plan rc=$?"
cap() { # name kernel-regex command...
  local n=$1 k=$2; shift 2
  timeout 420 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $O/r02f_$n "$@" > $O/r02f_ncufull_$n.log 2>&1; echo "full $n rc=$?"
  ncu -i $O/r02f_$n.ncu-rep --page raw --csv > $O/r02f_${n}_full.csv 2>/dev/null
}
cap plan_squad_kernel plan_squad python bench.py --steps 1 --warmup 1 --no-cpu --no-arena --lanes 1
cap plan_team_kernel plan_team python bench.py --steps 1 --warmup 1 --no-cpu --no-arena --lanes 1
rm -f $O/r02f_plan_squad_kernel.ncu-rep.tmp
ls -la $O/r02f_* | head -30
