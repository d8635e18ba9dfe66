#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "conv2d_slab" > gpurun_out/i_slab2d.log 2>&1; echo "rc=$?" >> gpurun_out/i_slab2d.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/i_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?" >> gpurun_out/i_bench.err
echo done
