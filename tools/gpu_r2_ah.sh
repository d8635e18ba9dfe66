#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_r2.py tests/test_gpu_render_parity.py -m gpu -q -x -k "mlp or render" > gpurun_out/ah_mlp_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ah_mlp_tests.log
timeout -k 10 300 python tools/run_render_branch.py > gpurun_out/ah_render.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/ah_render_launches.csv python tools/run_render_branch.py > gpurun_out/ah_render_ncu.log 2>&1
echo done
