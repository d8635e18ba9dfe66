#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 20 -c 12 -o gpurun_out/ad_wgrad python -c "
import os, sys, torch
os.environ['RCMVS_STRICT']='1'; sys.path.insert(0, '.'); sys.path.insert(0, 'rc-mvsnet_b200')
from rcmvs import train
print(train.bench_train_step(torch.device('cuda:0'), steps=1))
" > gpurun_out/ad_wgrad_ncu.log 2>&1
echo done
