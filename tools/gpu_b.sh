#!/bin/bash
# round-2 GPU call B: full -m gpu suite (no -x), collision workload A/B (staged vs ldg), launch list + ncu capture of the new kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -s ) > gpurun_out/r2b_gpu_suite.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_gpu_suite.log
grep -E "headline batch|mode \(i\)|mode \(ii\)|  query |^FAILED|^ERROR" gpurun_out/r2b_gpu_suite.log | head -60
timeout 300 python bench.py --workload collision --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_col_tma.log 2> gpurun_out/r2b_col_tma.err; echo "col tma rc=$?"
MPROS_C1_MODE=2phase_ldg timeout 300 python bench.py --workload collision --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_col_ldg.log 2> gpurun_out/r2b_col_ldg.err; echo "col ldg rc=$?"
python - <<'PY'
import json
for f in ["tma","ldg"]:
    try:
        d=json.loads(open("gpurun_out/r2b_col_%s.log"%f).read().strip().splitlines()[-1])
        print(f, "checks/s %.4g"%d["value"], "ms %.4f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/r2b_col_%s.err"%f).read()[-1500:])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:collision_poses_tma -c 1 -o gpurun_out/r02_collision_tma -f python bench.py --workload collision --steps 1 --warmup 3 --no-cpu > gpurun_out/r2b_ncu_col.log 2>&1; echo "ncu rc=$?"
