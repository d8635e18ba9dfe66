#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests -m gpu -q -x -k "softargmin or depth_index or regress" > gpurun_out/w_sa.log 2>&1; echo "rc=$?" >> gpurun_out/w_sa.log
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/w_tests.log
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; echo "bench rc=$?" >> gpurun_out/w_bench.err
for n in 3 4; do timeout -k 10 300 python bench.py --no-cpu-baseline --no-stock-torch --no-render --no-kernel-rooflines --in-flight $n > gpurun_out/w_bench_inflight$n.json 2>> gpurun_out/w_bench.err; done
echo done
