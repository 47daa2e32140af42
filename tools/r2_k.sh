#!/bin/bash
# round-2 GPU call K: CTA shapes of the squad kernel, lanes
O=gpurun_out; mkdir -p $O
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
for V in _q12 _q8 _q32 _q16x2; do
  MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200$V.so MPROS_DEBUG=1 timeout 600 python bench.py --lanes 1 --steps 1 --warmup 1 --no-cpu --no-arena > $O/r2k_dbg$V.json 2> $O/r2k_dbg$V.err; echo "dbg$V rc=$?"; grep "mpros" $O/r2k_dbg$V.err | tail -2
  MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200$V.so MPROS_SOLO_NODES=65536 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2k_plan$V.json 2> $O/r2k_plan$V.err; echo "plan$V rc=$?"; summ $O/r2k_plan$V.json
done
MPROS_SOLO_NODES=65536 timeout 600 python bench.py --steps 16 --warmup 8 --lanes 8 --no-cpu --no-arena > $O/r2k_plan_l8.json 2> $O/r2k_plan_l8.err; echo "plan l8 rc=$?"; summ $O/r2k_plan_l8.json
