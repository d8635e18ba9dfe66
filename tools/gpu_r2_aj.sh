#!/bin/bash
# last validation of the final tree: GPU suite, smoke, default bench without the CPU arms
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/aj_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/aj_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/aj_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/aj_smoke.log
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/aj_bench.json 2> gpurun_out/aj_bench.err; echo "bench rc=$?" >> gpurun_out/aj_bench.err
echo done
