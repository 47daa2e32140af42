#!/bin/bash
# round-2 GPU call N: squad kernel with grown workspaces
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_synthetic_fullsize.py -q -m gpu -s -x -k "solo or 16384 or batch or properties" ) > $O/r2n_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2n_tests.log
grep -E "FAILED|Error|error|differ" $O/r2n_tests.log | head -20
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
MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200_prof.so MPROS_DEBUG=1 timeout 600 python bench.py --lanes 1 --steps 1 --warmup 1 --no-cpu --no-arena > $O/r2n_prof.json 2> $O/r2n_prof.err; echo "prof rc=$?"; grep -B1 -A28 "squad profile" $O/r2n_prof.err | tail -32 | cut -c1-330
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2n_plan.json 2> $O/r2n_plan.err; echo "plan rc=$?"; summ $O/r2n_plan.json
MPROS_SOLO_NODES=65536 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2n_plan_n65536.json 2> $O/r2n_plan_n65536.err; echo "nodes 65536 rc=$?"; summ $O/r2n_plan_n65536.json
