#!/bin/bash
# round-2 final 2-GPU run: the multi-GPU paths over NCCL / CUDA IPC against the single-GPU results, weak-scaling bench line
O=gpurun_out; mkdir -p $O
nvidia-smi -L
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_banded_check.py 4096 ) > $O/r02f_mgpu_check_2gpu.log 2>&1; echo "mgpu rc=$?"
grep -v "^W\|Warning\|warn\|^\*\|OMP_NUM" $O/r02f_mgpu_check_2gpu.log | tail -8 | cut -c1-400
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu ) > $O/r02f_bench_weak2.log 2>&1; echo "weak rc=$?"; grep '^{' $O/r02f_bench_weak2.log | tail -1 | cut -c1-300
