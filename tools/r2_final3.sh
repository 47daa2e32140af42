#!/bin/bash
# round-2 final: the four kernel-class workloads at HEAD (with their CPU baselines) and the launch list of the rs workload
O=gpurun_out; mkdir -p $O
for W in collision expand rs preprocess; do
  timeout 240 python bench.py --workload $W > $O/r02f_bench_$W.json 2> $O/r02f_bench_$W.err; echo "$W rc=$?"
done
python - <<'PY'
import json
for w in ["collision","expand","rs","preprocess"]:
    try:
        d=json.loads(open("gpurun_out/r02f_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, d["metric"], "%.5g"%d["value"], "ms %.4g"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], "roof %.4f"%d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
    except Exception as e:
        print(w, "ERR", e); print(open("gpurun_out/r02f_bench_%s.err"%w).read()[-800:])
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 200 ncu --metrics $M --clock-control none -c 300 --csv --log-file $O/r02f_launches_rs_bench.csv python bench.py --workload rs --steps 2 --warmup 3 --no-cpu > $O/r02f_ncu_rs.log 2>&1; echo "launches rs rc=$?"
