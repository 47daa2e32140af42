#!/bin/bash
# round-2 GPU call F: solo (warp-per-query) pass + team pass. Parity first, then throughput of the register variants.
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_synthetic_fullsize.py tests/test_gpu_rs_expand_plan.py tests/test_gpu_case3_arena.py -q -m gpu -s -x -k "solo or 16384 or plan or batch or arena or case" ) > $O/r2f_tests.log 2>&1; echo "pytest rc=$?"; tail -8 $O/r2f_tests.log
grep -E "FAILED|Error|error|differ|nodes inserted" $O/r2f_tests.log | head -20
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
MPROS_DEBUG=1 timeout 600 python bench.py --lanes 1 --steps 2 --warmup 1 --no-cpu --no-arena > $O/r2f_dbg.json 2> $O/r2f_dbg.err; echo "dbg rc=$?"; grep "mpros" $O/r2f_dbg.err | tail -8
for V in "" _s32 _s16; do
  MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200$V.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2f_plan$V.json 2> $O/r2f_plan$V.err; echo "plan$V rc=$?"; summ $O/r2f_plan$V.json
done
MPROS_PLAN_MODE=team timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2f_plan_team.json 2> $O/r2f_plan_team.err; echo "team rc=$?"; summ $O/r2f_plan_team.json
for N in 16384 65536; do
  MPROS_SOLO_NODES=$N timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2f_plan_n$N.json 2> $O/r2f_plan_n$N.err; echo "nodes $N rc=$?"; summ $O/r2f_plan_n$N.json
done
