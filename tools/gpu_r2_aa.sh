#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_r2.py tests/test_gpu_render_parity.py -m gpu -q -k "render or volume or resize" > gpurun_out/aa_render_tests.log 2>&1; echo "rc=$?" >> gpurun_out/aa_render_tests.log
timeout -k 10 300 python tools/run_render_branch.py > gpurun_out/aa_render.log 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/aa_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/aa_tests.log
echo done
