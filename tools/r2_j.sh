#!/bin/bash
# round-2 GPU call J: squad pass hands long queries over WITH their state (continue, not restart); tail rule
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_synthetic_fullsize.py tests/test_gpu_rs_expand_plan.py tests/test_gpu_case3_arena.py -q -m gpu -s -x -k "solo or 16384 or batch or plan or case" ) > $O/r2j_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2j_tests.log
grep -E "FAILED|Error|error|differ|nodes inserted" $O/r2j_tests.log | head -20
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "plans/s %.0f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "single_batch_ms %.1f"%d["config"]["single_batch_ms"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "e2e %.0f"%d["e2e"]["value"])
except Exception as e:
    print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200_prof.so MPROS_DEBUG=1 timeout 600 python bench.py --lanes 1 --steps 1 --warmup 1 --no-cpu --no-arena > $O/r2j_prof.json 2> $O/r2j_prof.err; echo "prof rc=$?"; grep "mpros" $O/r2j_prof.err | grep -v "^\[mpros\]   [a-z]" | tail -6; grep -A18 "squad profile" $O/r2j_prof.err | tail -19
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2j_plan.json 2> $O/r2j_plan.err; echo "plan rc=$?"; summ $O/r2j_plan.json
timeout 600 python bench.py --steps 10 --warmup 3 --lanes 2 --no-cpu --no-arena > $O/r2j_plan_l2.json 2> $O/r2j_plan_l2.err; echo "plan l2 rc=$?"; summ $O/r2j_plan_l2.json
for N in 16384 65536; do
  MPROS_SOLO_NODES=$N timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2j_plan_n$N.json 2> $O/r2j_plan_n$N.err; echo "nodes $N rc=$?"; summ $O/r2j_plan_n$N.json
done
