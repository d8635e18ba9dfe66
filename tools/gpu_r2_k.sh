#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "s2_slab" > gpurun_out/k_s2.log 2>&1; echo "rc=$?" >> gpurun_out/k_s2.log
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/k_tests.log
timeout -k 10 600 python tools/parity_full.py > gpurun_out/k_parity_full.txt 2>&1; echo "rc=$?" >> gpurun_out/k_parity_full.txt
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?" >> gpurun_out/k_bench.err
for n in 1 3 4; do timeout -k 10 300 python bench.py --no-cpu-baseline --no-stock-torch --no-render --no-kernel-rooflines --in-flight $n > gpurun_out/k_bench_inflight$n.json 2>> gpurun_out/k_bench.err; done
echo done
