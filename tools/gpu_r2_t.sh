#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_tests.log
timeout -k 10 600 python tools/diag_tanks.py > gpurun_out/t_tanks.log 2>&1; echo "rc=$?" >> gpurun_out/t_tanks.log
timeout -k 10 600 python tools/parity_full.py > gpurun_out/t_parity_full.txt 2>&1; echo "rc=$?" >> gpurun_out/t_parity_full.txt
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?" >> gpurun_out/t_bench.err
echo done
