#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r2.py -m gpu -q -x -k "slab" > gpurun_out/b_slab.log 2>&1; echo "rc=$?" >> gpurun_out/b_slab.log
timeout 600 python tools/diag_tanks.py > gpurun_out/b_tanks.log 2>&1; echo "rc=$?" >> gpurun_out/b_tanks.log
timeout 600 python -m pytest tests/test_gpu_r2.py -m gpu -q -k "render_branch_full or second_device" > gpurun_out/b_rest.log 2>&1; echo "rc=$?" >> gpurun_out/b_rest.log
echo done
