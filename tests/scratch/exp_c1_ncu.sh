#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload collision --no-cpu --steps 10 > gpurun_out/bench_c1_two.log 2> gpurun_out/bench_c1_two.err; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01_collision.csv python bench.py --workload collision --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_c1_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:collision_poses_2phase -s 3 -c 1 -o gpurun_out/prof_collision_2phase_r01 -f python bench.py --workload collision --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_c1.log 2>&1; echo "ncu full rc=$?"
