#!/bin/bash
# round-2 GPU call E: phase profile + instruction-cache counters of the planner, team vs duo mode
mkdir -p gpurun_out
for M in team duo; do
  MPROS_PLAN_MODE=$M MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200_prof.so timeout 300 python tools/plan_profile.py 2048 > gpurun_out/r2e_prof_$M.log 2>&1; echo "prof $M rc=$?"
  MPROS_PLAN_MODE=$M timeout 600 ncu --metrics gpu__time_duration.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,sm__icc_requests_lookup_miss.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__inst_executed.sum --clock-control none -k regex:plan_ -c 2 --csv --log-file gpurun_out/r2e_ncu_$M.csv python tools/plan_profile.py 2048 > gpurun_out/r2e_ncu_$M.log 2>&1; echo "ncu $M rc=$?"
done
grep -A22 "profile:" gpurun_out/r2e_prof_team.log | head -24
grep -A22 "profile:" gpurun_out/r2e_prof_duo.log | head -24
python - <<'PY'
import csv
for m in ["team","duo"]:
    try:
        rows=[r for r in csv.reader(open("gpurun_out/r2e_ncu_%s.csv"%m)) if len(r)>10]
        hdr=rows[0]
        for r in rows[1:]:
            d=dict(zip(hdr,r))
            print(m, d.get("Kernel Name","")[:30], d.get("Metric Name"), d.get("Metric Value"))
    except Exception as e:
        print(m,"ERR",e)
PY
