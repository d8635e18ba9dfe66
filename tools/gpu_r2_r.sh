#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "o_slab or s2_slab" > gpurun_out/r_slab8.log 2>&1; echo "rc=$?" >> gpurun_out/r_slab8.log
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r_tests.log
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc=$?" >> gpurun_out/r_bench.err
echo done
