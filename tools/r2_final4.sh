#!/bin/bash
# round-2 final verification at HEAD (throughput policy in): GPU suite, smoke, the driver's bench command
O=gpurun_out; mkdir -p $O
( time timeout 400 python -m pytest tests -x -q -m gpu -s ) > $O/r02g_gpu_suite.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02g_gpu_suite.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02g_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02g_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02g_bench_plan.json 2> $O/r02g_bench_plan.err; echo "plan rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02g_bench_plan.json").read().strip().splitlines()[-1])
    print("plan", "%.5g"%d["value"], "ms %.4g"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], "single %.1f"%d["config"]["single_batch_ms"], "exp/s %.3g"%d["extra"]["node_expansions_per_sec"], "launches", d["gpu_launches"], "traffic", d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None, "arena", d["extra"].get("case3_arena",{}).get("plans_per_sec"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02g_bench_plan.err").read()[-1500:])
PY
