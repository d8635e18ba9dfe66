#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "sweep" > gpurun_out/af_sweep_test.log 2>&1; echo "rc=$?" >> gpurun_out/af_sweep_test.log
timeout -k 10 300 python tools/sweep_ab.py > gpurun_out/af_sweep_ab.log 2>&1; echo "rc=$?" >> gpurun_out/af_sweep_ab.log
echo done
