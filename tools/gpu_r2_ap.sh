#!/bin/bash
# N=2: graph-mode DDP training check (exit must be clean) + the bench command the driver runs at N=2 (short form)
set -u
mkdir -p gpurun_out
export H=512 W=640
GRAPH=1 timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/train_ddp_check.py > gpurun_out/ap_ddp_graph.log 2>&1; echo "rc=$?" >> gpurun_out/ap_ddp_graph.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-stock-torch --no-kernel-rooflines > gpurun_out/ap_bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/ap_bench_n2.log
echo done
