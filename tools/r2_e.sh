#!/bin/bash
# round-2 GPU call E (2 GPUs): the multi-GPU paths over NCCL / CUDA IPC against the single-GPU results
O=gpurun_out; mkdir -p $O
nvidia-smi -L
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_banded_check.py 4096 ) > $O/r2e_mgpu_check.log 2>&1; echo "mgpu rc=$?"
grep -v "^W\|Warning\|warn" $O/r2e_mgpu_check.log | tail -15
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu ) > $O/r2e_bench_weak2.log 2>&1; echo "weak rc=$?"; tail -1 $O/r2e_bench_weak2.log | cut -c1-400
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --scaling strong ) > $O/r2e_bench_strong2.log 2>&1; echo "strong rc=$?"; tail -1 $O/r2e_bench_strong2.log | cut -c1-400
