#!/bin/bash
# validation of the final tree: GPU suite, smoke, default bench (all arms), launch list of one training step
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/aq_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/aq_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/aq_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/aq_smoke.log
timeout -k 10 900 python bench.py > gpurun_out/aq_bench.json 2> gpurun_out/aq_bench.err; echo "bench rc=$?" >> gpurun_out/aq_bench.err
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 4700 --csv --log-file gpurun_out/aq_train_launches.csv python -c "
import os, sys, torch
os.environ['RCMVS_STRICT']='1'; sys.path.insert(0, '.'); sys.path.insert(0, 'rc-mvsnet_b200')
from rcmvs import train
print(train.bench_train_step(torch.device('cuda:0'), steps=2, graph=False))
" > gpurun_out/aq_train_ncu.log 2>&1
echo done
