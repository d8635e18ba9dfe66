#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/diag_tanks.py > gpurun_out/d_tanks.log 2>&1; echo "rc=$?" >> gpurun_out/d_tanks.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_r2.py::test_tanks_shaped_eval_matches_reference_golden > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/d_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?" >> gpurun_out/d_bench.err
echo done
