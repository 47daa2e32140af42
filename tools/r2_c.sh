#!/bin/bash
# round-2 GPU call C: whole GPU suite on the pooled workspaces / split Reeds-Shepp kernels / banded map, then the benches they touch
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests -x -q -m gpu -s ) > $O/r2c_tests.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r2c_tests.log
grep -E "rs_solve|rs_shot|FAILED|Error|error|differ" $O/r2c_tests.log | head -20
timeout 300 python bench.py --workload rs --steps 20 --warmup 5 --no-cpu > $O/r2c_rs.json 2> $O/r2c_rs.err; echo "rs rc=$?"
timeout 300 python bench.py --workload expand --steps 10 --warmup 3 --no-cpu > $O/r2c_expand.json 2> $O/r2c_expand.err; echo "expand rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-arena > $O/r2c_plan.json 2> $O/r2c_plan.err; echo "plan rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-arena --lanes 8 > $O/r2c_plan_l8.json 2> $O/r2c_plan_l8.err; echo "plan l8 rc=$?"
python - <<'PY'
import json
for f in ["rs","expand"]:
    try:
        d=json.loads(open("gpurun_out/r2c_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "%.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], d["e2e"].get("reference_format"))
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/r2c_%s.err"%f).read()[-1500:])
for f in ["plan","plan_l8"]:
    try:
        d=json.loads(open("gpurun_out/r2c_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "plans/s %.0f"%d["value"], "ms/step %.1f"%d["ms_per_step"], "single_batch_ms %.1f"%d["config"]["single_batch_ms"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "e2e %.0f"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/r2c_%s.err"%f).read()[-1500:])
PY
nvidia-smi --query-gpu=memory.used --format=csv
