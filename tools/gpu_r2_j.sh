#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "train or grad or wgrad or bwd or adam" > gpurun_out/j_train_tests.log 2>&1; echo "rc=$?" >> gpurun_out/j_train_tests.log
timeout 600 python tools/train_experiments.py > gpurun_out/j_train_ab.log 2>&1; echo "rc=$?" >> gpurun_out/j_train_ab.log
timeout 600 python tools/slab_experiments.py > gpurun_out/j_slab_exp.log 2>&1; echo "rc=$?" >> gpurun_out/j_slab_exp.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 4500 --csv --log-file gpurun_out/j_train_launches.csv python -c "
import os, sys, torch; os.environ['RCMVS_STRICT']='1'; sys.path.insert(0, '.')
sys.path.insert(0, 'rc-mvsnet_b200')
from rcmvs import train
print(train.bench_train_step(torch.device('cuda:0'), steps=1))
" > gpurun_out/j_train_ncu.log 2>&1
echo done
