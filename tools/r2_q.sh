#!/bin/bash
# round-2 GPU call Q: age limit of the throughput policy, eight batches in flight
O=gpurun_out; mkdir -p $O
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "plans/s %.0f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "e2e %.0f"%d["e2e"]["value"])
except Exception as e:
    print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-800:])
PY
}
for A in 12000 20000 50000; do
  MPROS_PLAN_AGE_LIMIT=$A timeout 120 python bench.py --steps 12 --warmup 8 --lanes 8 --no-cpu --no-arena > $O/r2q_a$A.json 2> $O/r2q_a$A.err; echo "age $A rc=$?"; summ $O/r2q_a$A.json
done
