#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 600 python tools/slab_experiments.py > gpurun_out/l_slab_exp.log 2>&1; echo "rc=$?" >> gpurun_out/l_slab_exp.log
timeout -k 10 600 python -m pytest tests/test_gpu_r2.py tests/test_gpu_render_parity.py -m gpu -q -x -k "render or slab" > gpurun_out/l_render_tests.log 2>&1; echo "rc=$?" >> gpurun_out/l_render_tests.log
timeout -k 10 900 python bench.py --no-cpu-baseline --no-stock-torch > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?" >> gpurun_out/l_bench.err
echo done
