#!/bin/bash
# round-2 GPU call H: squad kernel (W queries per CTA, CTA-wide heuristics). Parity, then throughput of the CTA shapes.
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_synthetic_fullsize.py -q -m gpu -s -x -k "solo or 16384 or batch" ) > $O/r2h_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2h_tests.log
grep -E "FAILED|Error|error|differ|nodes inserted" $O/r2h_tests.log | head -20
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
for V in "" _q12 _q16 _q8; do
  MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200$V.so MPROS_DEBUG=1 timeout 600 python bench.py --lanes 1 --steps 2 --warmup 1 --no-cpu --no-arena > $O/r2h_dbg$V.json 2> $O/r2h_dbg$V.err; echo "dbg$V rc=$?"; grep "mpros" $O/r2h_dbg$V.err | tail -2
done
for V in "" _q12; do
  MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200$V.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2h_plan$V.json 2> $O/r2h_plan$V.err; echo "plan$V rc=$?"; summ $O/r2h_plan$V.json
done
MPROS_SOLO_NODES=65536 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-arena > $O/r2h_plan_n65536.json 2> $O/r2h_plan_n65536.err; echo "nodes 65536 rc=$?"; summ $O/r2h_plan_n65536.json
