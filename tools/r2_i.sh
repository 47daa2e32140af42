#!/bin/bash
O=gpurun_out; mkdir -p $O
MPROS_LIB=$PWD/hybrid-astar-planner_b200/libmpros_b200_prof.so MPROS_DEBUG=1 timeout 600 python bench.py --lanes 1 --steps 1 --warmup 1 --no-cpu --no-arena > $O/r2i_prof.json 2> $O/r2i_prof.err; echo "prof rc=$?"; grep "mpros" $O/r2i_prof.err | tail -40
