#!/bin/bash
# round-2 GPU call R: batches in flight at the driver's step counts
O=gpurun_out; mkdir -p $O
for L in 8 6; do
  timeout 120 python bench.py --steps 20 --warmup 5 --lanes $L --no-cpu --no-arena > $O/r2r_l$L.json 2> $O/r2r_l$L.err; echo "lanes $L rc=$?"
  python - $O/r2r_l$L.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], "plans/s %.0f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "e2e %.0f"%d["e2e"]["value"])
PY
done
