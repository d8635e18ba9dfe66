#!/bin/bash
# ncu --set full of the 3x3x3 weight-gradient instantiations inside one training step
set -u
mkdir -p gpurun_out
timeout -k 10 500 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:wgrad_smem_kernel<.int.3, .int.3, .int.[12], .int.4' -s 30 -c 15 -o gpurun_out/as_wgrad3d -f python -c "
import os, sys, torch
os.environ['RCMVS_STRICT']='1'; sys.path.insert(0, '.'); sys.path.insert(0, 'rc-mvsnet_b200')
from rcmvs import train
print(train.bench_train_step(torch.device('cuda:0'), steps=1, graph=False))
" > gpurun_out/as_ncu.log 2>&1
echo done
