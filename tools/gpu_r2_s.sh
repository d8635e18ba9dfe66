#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "conv2d_slab or feature" > gpurun_out/s_slab8.log 2>&1; echo "rc=$?" >> gpurun_out/s_slab8.log
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s_tests.log
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?" >> gpurun_out/s_bench.err
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s_launches.csv \
  python bench.py --steps 2 --warmup 3 --reps 1 --no-cpu-baseline --no-stock-torch --no-render --no-kernel-rooflines > gpurun_out/s_bench_under_ncu.log 2>&1
echo done
