#!/bin/bash
# round-2: the driver's 2-GPU command at HEAD (eight batches in flight, throughput policy)
O=gpurun_out; mkdir -p $O
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 12 --warmup 5 --no-cpu ) > $O/r02g_bench_weak2.log 2>&1; echo "weak rc=$?"; grep '^{' $O/r02g_bench_weak2.log | tail -1 | cut -c1-200; grep -o '"per_rank": {[^}]*}' $O/r02g_bench_weak2.log | cut -c1-300
