#!/bin/bash
# round-2 GPU call L: ncu of the squad kernel
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plan_squad -c 1 -f -o $O/r2l_plan_squad_kernel python bench.py --lanes 1 --steps 1 --warmup 1 --no-cpu --no-arena > $O/r2l_ncu_squad.log 2>&1; echo "ncu rc=$?"
ncu -i $O/r2l_plan_squad_kernel.ncu-rep --page raw --csv > $O/r2l_plan_squad_kernel_full.csv 2>/dev/null
ncu -i $O/r2l_plan_squad_kernel.ncu-rep --page source --csv --print-source cuda,sass > $O/r2l_plan_squad_kernel_cs.csv 2>/dev/null
ls -la $O/r2l_*
