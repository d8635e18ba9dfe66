#!/bin/bash
# ncu --set full of the weight-gradient kernel inside one training step (3x3x3 stride-1 and stride-2 instantiations)
set -u
mkdir -p gpurun_out
timeout -k 10 500 ncu --set full --import-source on --clock-control none -k regex:wgrad_smem_kernel -s 60 -c 24 -o gpurun_out/ar_wgrad_smem -f python -c "
import os, sys, torch
os.environ['RCMVS_STRICT']='1'; sys.path.insert(0, '.'); sys.path.insert(0, 'rc-mvsnet_b200')
from rcmvs import train
print(train.bench_train_step(torch.device('cuda:0'), steps=1, graph=False))
" > gpurun_out/ar_ncu.log 2>&1
echo done
