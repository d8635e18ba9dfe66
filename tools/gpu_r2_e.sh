#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "vs_torch or adjoints or adam or training" > gpurun_out/e_train_tests.log 2>&1; echo "rc=$?" >> gpurun_out/e_train_tests.log
timeout 600 python -c "
import os,sys,json
sys.path.insert(0,'.'); sys.path.insert(0,'rc-mvsnet_b200'); os.environ['RCMVS_STRICT']='1'
import torch
from rcmvs import train
print(json.dumps(train.bench_train_step(torch.device('cuda:0'), steps=3)))
" > gpurun_out/e_train_bench.log 2>&1; echo "rc=$?" >> gpurun_out/e_train_bench.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/e_tests.log
echo done
