#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "sgemm or render_mlp_training or composite_point or rendering_consistency_net_training" > gpurun_out/f_render_train.log 2>&1; echo "rc=$?" >> gpurun_out/f_render_train.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1400 --csv --log-file gpurun_out/f_train_launches.csv python -c "
import os,sys
sys.path.insert(0,'.'); sys.path.insert(0,'rc-mvsnet_b200'); os.environ['RCMVS_STRICT']='1'
import torch
from rcmvs import train
print(train.bench_train_step(torch.device('cuda:0'), steps=1))
" > gpurun_out/f_train_ncu.log 2>&1
echo done
