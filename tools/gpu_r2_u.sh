#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "cuda_graph" > gpurun_out/u_graph.log 2>&1; echo "rc=$?" >> gpurun_out/u_graph.log
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?" >> gpurun_out/u_bench.err
timeout -k 10 600 python bench.py --no-cpu-baseline --no-stock-torch --no-render --no-kernel-rooflines --no-graphs > gpurun_out/u_bench_eager.json 2>> gpurun_out/u_bench.err
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/u_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/u_tests.log
echo done
