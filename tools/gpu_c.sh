#!/bin/bash
# round-2 GPU call C: synthetic full-size tests again + collision staging matrix
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_synthetic_fullsize.py tests/test_gpu_headless_facade.py -q -m gpu -s ) > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_tests.log
grep -E "^FAILED|^ERROR" gpurun_out/r2c_tests.log | head
for M in 2phase_s0 2phase_s1 2phase_s2 2phase_ldg; do
  MPROS_C1_MODE=$M timeout 300 python bench.py --workload collision --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c_col_$M.log 2> gpurun_out/r2c_col_$M.err; echo "$M rc=$?"
done
python - <<'PY'
import json
for f in ["2phase_s0","2phase_s1","2phase_s2","2phase_ldg"]:
    try:
        d=json.loads(open("gpurun_out/r2c_col_%s.log"%f).read().strip().splitlines()[-1])
        print(f, "checks/s %.4g"%d["value"], "ms %.4f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/r2c_col_%s.err"%f).read()[-1500:])
PY
