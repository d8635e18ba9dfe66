#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?" >> gpurun_out/c_bench.err
RCMVS_SWEEP=v2 RCMVS_CONV=tf32x3 timeout 600 python bench.py --no-cpu-baseline --no-render --no-stock-torch --reps 3 > gpurun_out/c_bench_r1cfg.json 2> gpurun_out/c_bench_r1cfg.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3d_slab_bf16 -c 1 -o gpurun_out/c_conv0_slab python tools/run_stage1_kernels.py 2 > gpurun_out/c_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:costvolume_var_v3 -c 1 -o gpurun_out/c_sweep_v3 python tools/run_stage1_kernels.py 2 > gpurun_out/c_ncu2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c_launches.csv \
  python bench.py --steps 2 --warmup 3 --reps 1 --no-cpu-baseline --no-stock-torch --no-render --no-kernel-rooflines > gpurun_out/c_bench_under_ncu.log 2>&1
echo done
