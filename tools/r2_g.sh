#!/bin/bash
# round-2 GPU call G: ncu of the solo kernel
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plan_solo -c 1 -f -o $O/r2g_plan_solo_kernel python bench.py --lanes 1 --steps 1 --warmup 1 --no-cpu --no-arena > $O/r2g_ncu_solo.log 2>&1; echo "ncu rc=$?"
ncu -i $O/r2g_plan_solo_kernel.ncu-rep --page raw --csv > $O/r2g_plan_solo_kernel_full.csv 2>/dev/null
ncu -i $O/r2g_plan_solo_kernel.ncu-rep --page source --csv > $O/r2g_plan_solo_kernel_source.csv 2>/dev/null
ls -la $O/r2g_*
