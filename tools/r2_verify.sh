#!/bin/bash
# round-2 verification run (developer run, not a test): what the driver runs at round end, every bench workload,
# launch lists and one `ncu --set full` capture per hot kernel (each after its plain command exited 0).
# Outputs land in gpurun_out/r02v_*; the summaries judged are copied into profiles/ by hand afterwards.
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -x -q -m gpu -s ) > $O/r02v_gpu_suite.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02v_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02v_smoke.log
timeout 900 python bench.py > $O/r02v_bench_plan.json 2> $O/r02v_bench_plan.err; echo "plan rc=$?"
for W in collision expand rs preprocess; do
  timeout 400 python bench.py --workload $W > $O/r02v_bench_$W.json 2> $O/r02v_bench_$W.err; echo "$W rc=$?"
done
python - <<'PY'
import json
for w in ["plan","collision","expand","rs","preprocess"]:
    try:
        d=json.loads(open("gpurun_out/r02v_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, d["metric"], "%.5g"%d["value"], "ms %.4g"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], "roof %.4f"%d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
    except Exception as e:
        print(w, "ERR", e); print(open("gpurun_out/r02v_bench_%s.err"%w).read()[-1200:])
PY
# launch lists (gpu__time_duration + DRAM bytes), same commands with few steps
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none -c 200 --csv --log-file $O/r02v_launches_plan_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-arena > $O/r02v_ncu_plan.log 2>&1; echo "launches plan rc=$?"
for W in collision preprocess rs expand; do
  timeout 400 ncu --metrics $M --clock-control none -c 300 --csv --log-file $O/r02v_launches_${W}_bench.csv python bench.py --workload $W --steps 2 --warmup 3 --no-cpu > $O/r02v_ncu_$W.log 2>&1; echo "launches $W rc=$?"
done
# full captures, one launch each
cap() { # name kernel-regex command...
  local n=$1 k=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $O/r02v_$n "$@" > $O/r02v_ncufull_$n.log 2>&1; echo "full $n rc=$?"
  ncu -i $O/r02v_$n.ncu-rep --page raw --csv > $O/r02v_${n}_full.csv 2>/dev/null
}
cap collision_poses_tma_kernel collision_poses_tma python bench.py --workload collision --steps 1 --warmup 3 --no-cpu
cap rs_batch_kernel rs_batch python bench.py --workload rs --steps 1 --warmup 3 --no-cpu
cap expand_batch_kernel expand_batch python bench.py --workload expand --steps 1 --warmup 3 --no-cpu
cap goal_bfs_cluster_kernel goal_bfs_cluster python bench.py --workload preprocess --steps 1 --warmup 3 --no-cpu
cap plan_team_kernel plan_team python bench.py --steps 1 --warmup 3 --no-cpu --no-arena --lanes 1
ls -la $O | head -60
