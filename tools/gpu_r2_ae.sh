#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "forward_backward or training or wgrad or adam or bwd" > gpurun_out/ae_train_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ae_train_tests.log
timeout -k 10 600 python tools/train_experiments.py > gpurun_out/ae_train_ab.log 2>&1; echo "rc=$?" >> gpurun_out/ae_train_ab.log
echo done
