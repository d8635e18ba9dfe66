#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "conv3d_slab_bf16x3" > gpurun_out/x_slab.log 2>&1; echo "rc=$?" >> gpurun_out/x_slab.log
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/x_tests.log
timeout -k 10 600 python tools/diag_tanks.py > gpurun_out/x_tanks.log 2>&1
timeout -k 10 600 python tools/parity_full.py > gpurun_out/x_parity_full.txt 2>&1
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$?" >> gpurun_out/x_bench.err
echo done
