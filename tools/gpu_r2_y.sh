#!/bin/bash
# ncu --set full captures of the new TMA kernels (first launch of each = stage 1) + a final launch list
set -u
mkdir -p gpurun_out
for k in convT3d_slab_bf16_kernel conv3d_s2_slab_bf16_kernel conv2d_s2k5_slab_bf16_kernel conv3d_o_slab_bf16_kernel; do
  timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/y_$k python tools/run_stage1_full.py 1 > gpurun_out/y_ncu_$k.log 2>&1
done
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/y_launches.csv \
  python bench.py --steps 2 --warmup 3 --reps 1 --no-graphs --no-cpu-baseline --no-stock-torch --no-render --no-kernel-rooflines > gpurun_out/y_bench_under_ncu.log 2>&1
echo done
