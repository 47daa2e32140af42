#!/bin/bash
# round-2 GPU call D: whole GPU suite again (pool sized by demand, shared tickets), default bench, ncu of the two Reeds-Shepp kernels
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests -q -m gpu -s ) > $O/r2d_tests.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r2d_tests.log
grep -E "FAILED|Error|error|differ" $O/r2d_tests.log | head -20
timeout 900 python bench.py > $O/r2d_plan.json 2> $O/r2d_plan.err; echo "plan rc=$?"
python - <<'PY'
import json
for f in ["plan"]:
    try:
        d=json.loads(open("gpurun_out/r2d_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "plans/s %.0f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "single_batch_ms %.1f"%d["config"]["single_batch_ms"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "e2e %.0f"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/r2d_%s.err"%f).read()[-1500:])
PY
for K in rs_solve rs_traj; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o $O/r2d_${K}_kernel python bench.py --workload rs --steps 1 --warmup 3 --no-cpu > $O/r2d_ncu_$K.log 2>&1; echo "ncu $K rc=$?"
  ncu -i $O/r2d_${K}_kernel.ncu-rep --page raw --csv > $O/r2d_${K}_kernel_full.csv 2>/dev/null
done
