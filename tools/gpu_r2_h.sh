#!/bin/bash
# 2-GPU call: multi-device guard test, DDP training check (NCCL), bench at N=2 under torchrun
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/h_gpus.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_r2.py -m gpu -q -k "second_device" > gpurun_out/h_multidev.log 2>&1; echo "rc=$?" >> gpurun_out/h_multidev.log
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_ddp_check.py > gpurun_out/h_ddp_check.log 2>&1; echo "rc=$?" >> gpurun_out/h_ddp_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/h_bench_n2.json 2> gpurun_out/h_bench_n2.err; echo "rc=$?" >> gpurun_out/h_bench_n2.err
echo done
