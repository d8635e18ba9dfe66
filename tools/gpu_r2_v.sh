#!/bin/bash
# 4-GPU call: bench at N=4 under torchrun (inference headline + train_step with SyncBN + flat-gradient all-reduce over NCCL)
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/v_gpus.txt 2>&1
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-stock-torch > gpurun_out/v_bench_n4.json 2> gpurun_out/v_bench_n4.err; echo "rc=$?" >> gpurun_out/v_bench_n4.err
NCCL_DEBUG=WARN timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 tools/train_ddp_check.py > gpurun_out/v_ddp_check.log 2>&1; echo "rc=$?" >> gpurun_out/v_ddp_check.log
echo done
