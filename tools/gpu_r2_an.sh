#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 7000 -c 2400 --csv --log-file gpurun_out/an_train_launches.csv python -c "
import os, sys, torch
os.environ['RCMVS_STRICT']='1'; sys.path.insert(0, '.'); sys.path.insert(0, 'rc-mvsnet_b200')
from rcmvs import train
print(train.bench_train_step(torch.device('cuda:0'), steps=1, graph=False))
" > gpurun_out/an_train_ncu.log 2>&1
echo done
