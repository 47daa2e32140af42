#!/bin/bash
# round-2 GPU call P: throughput policy (migration list + adoption + age limit). Tight timeouts.
O=gpurun_out; mkdir -p $O
timeout 60 python tools/sanitize_plan.py 2>&1 | tail -6
( time timeout 200 python -m pytest tests/test_gpu_synthetic_fullsize.py -q -m gpu -s -x -k "solo" ) > $O/r2p_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2p_tests.log; grep -E "FAILED|Error|differ" $O/r2p_tests.log | head -5
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "plans/s %.0f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "single_batch_ms %.1f"%d["config"]["single_batch_ms"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "e2e %.0f"%d["e2e"]["value"])
except Exception as e:
    print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1200:])
PY
}
MPROS_PLAN_POLICY=throughput MPROS_DEBUG=1 timeout 120 python bench.py --lanes 1 --steps 1 --warmup 1 --no-cpu --no-arena > $O/r2p_dbg.json 2> $O/r2p_dbg.err; echo "dbg rc=$?"; grep "mpros" $O/r2p_dbg.err | tail -2 | cut -c1-330
timeout 150 python bench.py --steps 12 --warmup 4 --no-cpu --no-arena > $O/r2p_plan.json 2> $O/r2p_plan.err; echo "plan rc=$?"; summ $O/r2p_plan.json
timeout 150 python bench.py --steps 16 --warmup 8 --lanes 8 --no-cpu --no-arena > $O/r2p_plan_l8.json 2> $O/r2p_plan_l8.err; echo "plan l8 rc=$?"; summ $O/r2p_plan_l8.json
