#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "forward_backward or training or train_step or wgrad or adam or bwd" > gpurun_out/am_train_tests.log 2>&1; echo "rc=$?" >> gpurun_out/am_train_tests.log
timeout -k 10 600 python tools/train_experiments.py > gpurun_out/am_train_ab.log 2>&1; echo "rc=$?" >> gpurun_out/am_train_ab.log
[ "${1:-}" = "ncu" ] && timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 4500 --csv --log-file gpurun_out/am_train_launches.csv python -c "
import os, sys, torch
os.environ['RCMVS_STRICT']='1'; sys.path.insert(0, '.'); sys.path.insert(0, 'rc-mvsnet_b200')
from rcmvs import train
print(train.bench_train_step(torch.device('cuda:0'), steps=1, graph=False))
" > gpurun_out/am_train_ncu.log 2>&1
echo done
