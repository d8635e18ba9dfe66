#!/bin/bash
set -u
mkdir -p gpuruo_out
timeout -k 10 300 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "transpose3d_slab" > gpuruo_out/o_slab8.log 2>&1; echo "rc=$?" >> gpuruo_out/o_slab8.log
timeout -k 10 900 python -m pytest tests -m gpu -q > gpuruo_out/o_tests.log 2>&1; echo "tests rc=$?" >> gpuruo_out/o_tests.log
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpuruo_out/o_bench.json 2> gpuruo_out/o_bench.err; echo "bench rc=$?" >> gpuruo_out/o_bench.err
echo done
