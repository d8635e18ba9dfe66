#!/bin/bash
# N=2: TrainStep with SyncBatchNorm, eager vs CUDA-graph replay (NCCL all-reduces of the statistics captured)
set -u
mkdir -p gpurun_out
export H=512 W=640
GRAPH=0 timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/train_ddp_check.py > gpurun_out/ao_ddp_eager.log 2>&1; echo "rc=$?" >> gpurun_out/ao_ddp_eager.log
GRAPH=1 timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/train_ddp_check.py > gpurun_out/ao_ddp_graph.log 2>&1; echo "rc=$?" >> gpurun_out/ao_ddp_graph.log
echo done
