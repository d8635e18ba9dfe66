#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.log
timeout 900 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?" >> gpurun_out/g_bench.err
echo done
